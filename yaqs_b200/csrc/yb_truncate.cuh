// truncate() of the reference on a descending spectrum -- the keep decision of every SVD on the path (core/linalg/svd_utils.py:22-104,
// called from split_two_site, core/methods/decompositions.py:105-185, and the SVD gauge shifts, core/data_structures/mps.py:747-764).
// A header of its own so that the SAME source compiles for the host (YB_HOST_EMU; tests/host_emu/truncate_emu.cpp), where it is pinned
// against the known answers of the reference's own tests (tests/core/linalg/test_svd_utils.py) and the oracle's restatement.
#pragma once
#include <math.h>

#ifdef YB_HOST_EMU
#define YB_TRUNC_FN static inline
#define YB_TRUNC_MIN(a, b) ((a) < (b) ? (a) : (b))
#define YB_TRUNC_MAX(a, b) ((a) > (b) ? (a) : (b))
#else
#define YB_TRUNC_FN __device__
#define YB_TRUNC_MIN(a, b) min(a, b)
#define YB_TRUNC_MAX(a, b) max(a, b)
#endif

// (serial: the discarded-weight mode accumulates from the smallest value upward, and that summation order is part of the keep decision)
YB_TRUNC_FN int truncate_keep(const double* s, int n, int mode, double threshold, int max_bond, int min_keep) {
    if (n == 0) return 0;
    int keep = n;
    if (mode == 2) {                     // hard_cutoff
        keep = 0;
        for (int i = 0; i < n; ++i) keep += s[i] > threshold;
    } else if (mode == 1) {              // relative
        keep = 0;
        if (s[0] > 0.0) for (int i = 0; i < n; ++i) keep += (s[i] / s[0]) >= threshold;
    } else if (mode == 0) {              // discarded_weight
        double acc = 0.0;
        for (int idx = 0; idx < n; ++idx) {
            const double v = s[n - 1 - idx];
            acc += v * v;
            if (acc >= threshold) { keep = YB_TRUNC_MAX(n - idx, min_keep); break; }
        }
    } else {                             // relative_discarded_weight
        if (s[0] <= 0.0) keep = 0;
        else {
            double total = 0.0;
            for (int i = 0; i < n; ++i) { const double x = s[i] / s[0]; total += x * x; }
            double acc = 0.0;
            for (int idx = 0; idx < n; ++idx) {
                const double x = s[n - 1 - idx] / s[0];
                const double cand = acc + x * x;
                if (cand / total <= threshold) { acc = cand; keep = n - idx - 1; } else break;
            }
        }
    }
    if (max_bond > 0) keep = YB_TRUNC_MIN(keep, max_bond);
    keep = YB_TRUNC_MAX(keep, min_keep);
    return YB_TRUNC_MIN(keep, n);
}
