// EXPERIMENT (not in the default build; enable with build.py --variant ca TB_EXP_CA): pass C of one group fused
// with pass A of the next group on the same overlap stream.  Bit-identical to the separate passes, measured slower on
// the headline config (49.5 k fused / 48.3 k fused + async x tile vs 50.4 k matvecs/s, profiles/experiments_r02.md).
#pragma once
#include "../fft_kernels.cuh"

namespace tb {

// Last pass of one group fused with the first pass of the NEXT group that uses the same workspace (same overlap
// stream): the CTA that owns a column tile of the workspace turns it into y of the finished group (pass C) and then
// refills it from x of the next group (pass A).  One launch and one CTA ramp instead of two per column pair, and --
// STAGE -- the next group's x tile is fetched asynchronously (LDGSTS into a private staging area) while the inverse
// transform of the finished group runs, so the HBM latency of pass A disappears behind pass C.
// grid.x = (N2/TA) * max(pc.nfft, pa.nfft) ; HALF as in k_colsA / k_colsC (both sides padded, every Toeplitz / Hankel product).
template <typename R, int LOG2L1, int TA, bool STAGE>
constexpr int ca_smem_bytes() {
    return (int)sizeof(cx<R>) * (1 << LOG2L1) * TA + (STAGE ? (int)sizeof(R) * (1 << LOG2L1) * TA : 0);
}
template <typename R, int LOG2L1, int TA, bool HALF, bool STAGE>
__global__ void __launch_bounds__((1 << (LOG2L1 - le_of<R>())) * TA, min_ctas<R>((1 << (LOG2L1 - le_of<R>())) * TA, ca_smem_bytes<R, LOG2L1, TA, STAGE>()))
k_colsCA(const ConvArgs<R> pc, const ConvArgs<R> pa) {
    using F = SlotFFT<R, LOG2L1>;
    constexpr int NT = F::NT, E = F::E, EIN = HALF ? E / 2 : E, NTHR = NT * TA;
    static_assert(!STAGE || HALF, "the staging area is sized for half-padded input");
    extern __shared__ __align__(16) unsigned char smraw[];
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const int l2 = pc.log2n2;
    const unsigned tiles = (1u << l2) / TA;
    const long long f = blockIdx.x / tiles;
    const int j2 = (int)(blockIdx.x % tiles) * TA + q;
    ColMap<TA> map{q};
    const bool doC = f < pc.nfft, doA = f < pa.nfft;      // uniform per CTA
    // staging: [2 * EIN][NTHR] reals behind the FFT region; a thread only ever touches its own column of it
    R* stg = reinterpret_cast<R*>(smraw + sizeof(cx<R>) * (size_t)(1 << LOG2L1) * TA) + threadIdx.x;
    if constexpr (STAGE) {
        if (doA) {
            const InView<R> in(pa, f);
#pragma unroll
            for (int c = 0; c < EIN; ++c) in.stage_real(((t + c * NT) << l2) + j2, stg + (2 * c) * NTHR, stg + (2 * c + 1) * NTHR);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    cx<R> a[E];
    if (doC) {
        const cx<R>* in = pc.scratch + (f << (l2 + LOG2L1)) + j2;
        {
            cx<R> w[E];
            level_twiddles<R, F>(pc, t, j2, w);
#pragma unroll
            for (int i = 0; i < E; ++i) a[i] = cmulc(ld_stream(in + ((long long)F::pos_of_reg(i, t) << l2)), w[i]);
        }
        F::inverse(a, t, sm, map, pc.tw);
        const OutView<R> out(pc, f);
#pragma unroll
        for (int j = 0; j < EIN; ++j) out.store(((t + j * NT) << l2) + j2, a[j]);
    }
    __syncthreads();           // the inverse transform's last shared-memory reads precede the forward transform's first stores
    if (doA) {
        if constexpr (STAGE) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int c = 0; c < EIN; ++c) a[c] = mk<R>(stg[(2 * c) * NTHR], stg[(2 * c + 1) * NTHR]);
        } else {
            const InView<R> in(pa, f);
#pragma unroll
            for (int c = 0; c < EIN; ++c) a[c] = in.load(((t + c * NT) << l2) + j2);
        }
        F::template forward<ColMap<TA>, false, HALF>(a, t, sm, map, pa.tw);
        cx<R>* out = pa.scratch + (f << (l2 + LOG2L1)) + j2;
        cx<R> w[E];
        level_twiddles<R, F>(pa, t, j2, w);
#pragma unroll
        for (int i = 0; i < E; ++i) out[(long long)F::pos_of_reg(i, t) << l2] = cmul(a[i], w[i]);
    }
}

}  // namespace tb
