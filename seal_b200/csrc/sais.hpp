// Suffix array construction by induced sorting (SA-IS, Nong/Zhang/Chan 2009) for integer
// alphabets.  Host-side index construction only — replaces what the reference gets from sdsl's
// qsufsort (sdsl/construct_sa.hpp:162-166).  The text must end with a unique smallest symbol.
#pragma once
#include <cstdint>
#include <vector>

namespace sealb200 {

template <typename Sym, typename Idx>
class SaIs {
  public:
    // s[0..n), symbols in [0,K); s[n-1] is the unique minimum.  SA receives n entries.
    static void run(const Sym* s, Idx* SA, Idx n, Idx K) {
        if (n == 1) { SA[0] = 0; return; }
        std::vector<bool> stype(static_cast<size_t>(n));
        stype[n - 1] = true;
        for (Idx i = n - 2; i >= 0; --i)
            stype[i] = (s[i] < s[i + 1]) || (s[i] == s[i + 1] && stype[i + 1]);
        auto is_lms = [&](Idx i) { return i > 0 && stype[i] && !stype[i - 1]; };

        std::vector<Idx> bkt(static_cast<size_t>(K));
        auto buckets = [&](bool ends) {
            std::fill(bkt.begin(), bkt.end(), Idx(0));
            for (Idx i = 0; i < n; ++i) bkt[s[i]]++;
            Idx sum = 0;
            for (Idx c = 0; c < K; ++c) { sum += bkt[c]; bkt[c] = ends ? sum : sum - bkt[c]; }
        };
        auto induce = [&]() {
            buckets(false);
            for (Idx i = 0; i < n; ++i) {
                Idx j = SA[i] - 1;
                if (SA[i] > 0 && !stype[j]) SA[bkt[s[j]]++] = j;
            }
            buckets(true);
            for (Idx i = n - 1; i >= 0; --i) {
                Idx j = SA[i] - 1;
                if (SA[i] > 0 && stype[j]) SA[--bkt[s[j]]] = j;
            }
        };

        // 1. sort LMS substrings
        for (Idx i = 0; i < n; ++i) SA[i] = -1;
        buckets(true);
        for (Idx i = 1; i < n; ++i) if (is_lms(i)) SA[--bkt[s[i]]] = i;
        induce();

        // 2. name them
        Idx n1 = 0;
        for (Idx i = 0; i < n; ++i) if (is_lms(SA[i])) SA[n1++] = SA[i];
        for (Idx i = n1; i < n; ++i) SA[i] = -1;
        Idx name = 0, prev = -1;
        for (Idx i = 0; i < n1; ++i) {
            Idx pos = SA[i];
            bool diff = false;
            if (prev < 0) diff = true;
            else for (Idx d = 0;; ++d) {
                if (s[pos + d] != s[prev + d] || stype[pos + d] != stype[prev + d]) { diff = true; break; }
                if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
            }
            if (diff) { ++name; prev = pos; }
            SA[n1 + pos / 2] = name - 1;
        }
        for (Idx i = n - 1, j = n - 1; i >= n1; --i) if (SA[i] >= 0) SA[j--] = SA[i];

        // 3. sort the reduced problem
        Idx* SA1 = SA;
        Idx* s1 = SA + n - n1;
        if (name < n1) SaIs<Idx, Idx>::run(s1, SA1, n1, name);
        else for (Idx i = 0; i < n1; ++i) SA1[s1[i]] = i;

        // 4. induce the final order
        buckets(true);
        for (Idx i = 1, j = 0; i < n; ++i) if (is_lms(i)) s1[j++] = i;
        for (Idx i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
        for (Idx i = n1; i < n; ++i) SA[i] = -1;
        for (Idx i = n1 - 1; i >= 0; --i) {
            Idx j = SA[i]; SA[i] = -1;
            SA[--bkt[s[j]]] = j;
        }
        induce();
    }
};

}  // namespace sealb200
