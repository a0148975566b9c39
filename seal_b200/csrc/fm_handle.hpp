// Internal (non-ABI) access to the device view of an sealfm_t for the other translation units.
#pragma once
#include "../../include/sealfm.h"
#include "fm_device.cuh"

namespace sealb200 {
FmView sealfm_view(const sealfm_t* h);   // defined in fm_kernels.cu; throws ApiError if not on a device
}
