// distance.cuh — the DISTANCE CASE expression as one device function (shared by distance.cu / nearest.cu).
// reference: abdenlab/giql src/giql/expanders/_distance.py:67-117
#pragma once
#include "common.cuh"

// a, b on the SAME chromosome, canonical coordinates.  Returns false for SQL NULL.
__host__ __device__ __forceinline__ bool giqlb_distance_dev(long long as, long long ae, uint8_t sa, long long bs, long long be,
                                                           uint8_t sb, bool stranded, bool signed_, long long* out) {
    if (stranded) {
        // _distance.py:93-95 / :107-109 — NULL, '.' and '?' strands give NULL
        if (sa == GIQLB_STRAND_NULL || sb == GIQLB_STRAND_NULL) return false;
        if (sa == GIQLB_STRAND_DOT || sa == GIQLB_STRAND_QMARK) return false;
        if (sb == GIQLB_STRAND_DOT || sb == GIQLB_STRAND_QMARK) return false;
    }
    if (as < be && ae > bs) {            // _distance.py:74
        *out = 0;
        return true;
    }
    const bool a_minus = stranded && sa == GIQLB_STRAND_MINUS;
    if (ae <= bs) {                      // downstream, _distance.py:75 / :97-99
        long long gap = bs - ae + 1;
        *out = a_minus ? -gap : gap;
        return true;
    }
    long long gap = as - be + 1;         // upstream
    if (!stranded) *out = signed_ ? -gap : gap;                 // _distance.py:76 / :83
    else if (signed_) *out = a_minus ? gap : -gap;              // _distance.py:101-102
    else *out = a_minus ? -gap : gap;                           // _distance.py:115-116
    return true;
}
