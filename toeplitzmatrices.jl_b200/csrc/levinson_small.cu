// Levinson / Durbin at small orders: groups of 4 / 8 / 16 lanes per system (levinson_kernels.cuh).
#include "levinson_kernels.cuh"

namespace tb {

template <typename R, bool DURBIN>
cudaError_t launch_levinson_small(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                  const R* diag, cudaStream_t s) {
    const int wpc = 4;
#define TB_LEV_SMALL(SV, PV)                                                                                        \
    {                                                                                                               \
        const unsigned grid = (unsigned)((nsys + wpc * (32 / PV) - 1) / (wpc * (32 / PV)));                        \
        const size_t smem = (size_t)wpc * (32 / PV) * 2 * SV * (PV * LB) * sizeof(R);                               \
        k_levinson_warp<R, SV, DURBIN, PV><<<grid, wpc * 32, smem, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag);    \
    }
        // smallest group of lanes that still holds a system in S <= 4 register blocks per lane
        if (n <= 16) TB_LEV_SMALL(1, 4)
        else if (n <= 32) TB_LEV_SMALL(2, 4)
        else if (n <= 48) TB_LEV_SMALL(3, 4)
        else if (n <= 64) TB_LEV_SMALL(4, 4)
        else if (n <= 128 && !DURBIN) TB_LEV_SMALL(2, 16)      // measured: Levinson 102 M (16 lanes) vs 93 M (8 lanes) at n = 128
        else if (n <= 96) TB_LEV_SMALL(3, 8)
        else if (n <= 128) TB_LEV_SMALL(4, 8)
        else if (n <= 192) TB_LEV_SMALL(3, 16)
        else TB_LEV_SMALL(4, 16)
#undef TB_LEV_SMALL
    return cudaGetLastError();
}

#define TB_INST(R, D) template cudaError_t launch_levinson_small<R, D>(int64_t, int64_t, const R*, int64_t, const R*, int64_t, R*, int64_t, const R*, cudaStream_t);
TB_INST(float, false) TB_INST(float, true) TB_INST(double, false) TB_INST(double, true)
#undef TB_INST

}  // namespace tb
