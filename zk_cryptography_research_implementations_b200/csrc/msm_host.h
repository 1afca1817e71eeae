// msm_host.h -- the host finish of a multi-scalar multiplication: from the per-window plane sums the device returns to the
// affine result.  Pure host code (host_curve.h); kzg.cu calls it after the last kernel, tests/host_emu/emu_msm.cpp on the CPU.
#pragma once
#include <vector>
#include "host_curve.h"
#include "msm_plan.cuh"

namespace zk {

// planes: W x (nb + 1) points of one group: for window w, planes[w * (nb + 1) + p] = P_{w,p} for p < nb (the sum of the chunk
// sums run_t over the chunks t whose bit p is set) and planes[w * (nb + 1) + nb] = A_w (the sum of the chunks' weighted sums).
//   window_w = A_w + S sum_p 2^p P_{w,p};   result = sum_w 2^(c w) window_w
// i.e. a sum over bit positions: A_w sits at c w and P_{w,p} at c w + log2(S) + p < c (w + 1).  One pass from the top bit down:
// a doubling per position, an addition per term, one inversion for the affine form.
inline HG1Affine msm_combine_planes(const HG1Xyzz* planes, const MsmPlan& pl, uint32_t nb, uint32_t log_s) {
    std::vector<const HG1Xyzz*> at((size_t)pl.W * pl.c, nullptr), at2((size_t)pl.W * pl.c, nullptr);
    for (int w = 0; w < pl.W; ++w) {
        const HG1Xyzz* base = planes + (size_t)w * (nb + 1);
        at[(size_t)w * pl.c] = base + nb;
        for (uint32_t p = 0; p < nb; ++p) {
            const size_t pos = (size_t)w * pl.c + log_s + p;
            (at[pos] ? at2[pos] : at[pos]) = base + p;
        }
    }
    HG1Xyzz acc = HostG1::infinity();
    for (size_t pos = at.size(); pos-- > 0;) {
        acc = HostG1::dbl(acc);
        if (at[pos]) acc = HostG1::add(acc, *at[pos]);
        if (at2[pos]) acc = HostG1::add(acc, *at2[pos]);
    }
    return HostG1::to_affine(acc);
}

}  // namespace zk
