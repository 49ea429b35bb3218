// mlp_common.cuh — offsets of the flat fp32 parameter vector (include/tmla.h: policy.parameters() order)
#pragma once
#include <stdint.h>

struct MlpOffsets {
    int64_t w1[2], b1[2], w2[2], b2[2], wh[2], bh[2], total;
    int nout[2];
};
static inline MlpOffsets mlp_offsets(int D, int A) {
    MlpOffsets o;
    int64_t p = 0;
    for (int t = 0; t < 2; ++t) {
        o.w1[t] = p; p += (int64_t)256 * D;
        o.b1[t] = p; p += 256;
        o.w2[t] = p; p += (int64_t)256 * 256;
        o.b2[t] = p; p += 256;
    }
    o.wh[0] = p; p += (int64_t)A * 256;
    o.bh[0] = p; p += A;
    o.wh[1] = p; p += 256;
    o.bh[1] = p; p += 1;
    o.total = p;
    o.nout[0] = A; o.nout[1] = 1;
    return o;
}

