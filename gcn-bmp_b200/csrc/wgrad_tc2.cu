// wgrad_tc2.cu -- parameter-gradient contraction over the PANEL stash (BMP_MODE_BF16, stash v2).
//
// The tcgen05 encoder kernels dump their bf16 operand panels ([128 rows][64 cols], 128-byte swizzle)
// verbatim to global memory (cp.async.bulk stores).  A 64-row half of such a panel is, byte for byte, an
// MN-major UMMA operand block [64 k-rows][64 columns], so C += A^T B needs no conversion at all here:
// one lane streams 8 KB blocks with cp.async.bulk into a 3-stage ring, one lane issues
// M=128 x N x 16 UMMAs into a TMEM accumulator that lives for the CTA's whole row range (split over CTAs),
// four warps add the tile into C with fp32 atomics.  Bias gradients (column sums of A) come from one extra
// N=16 UMMA against a block of ones.
#include <cstdlib>
#include "tc_common.cuh"

namespace bmp {
namespace w2 {
using namespace tc;

constexpr int BLK = 64 * 128;          // 8 KB operand block
constexpr int STAGES = 3;
constexpr int NTH = 192;               // warps 0-3 epilogue, 4 TMA, 5 MMA

struct Args {
    const uint8_t *A, *B;              // panel arrays: [t][tile][ppt panels][16 KB]
    int a_ppt, b_ppt;                  // panels per tile in A / B
    int a_panel[2];                    // the two 64-column blocks of A forming the M = 128 tile (-1: absent)
    int b_panel[4];                    // N/64 blocks of B
    int nb;                            // N / 64
    float *C[2];                       // per A block: 64 rows of C (row-major, ldc)
    int ldc;
    float *bias[2];                    // per A block: column sums target (stride bias_stride) or NULL
    int bias_stride;
    int t0, t1, n_tiles;               // steps [t0, t1], tiles per step
    long chunks_per_cta;               // 64-row chunks per CTA
};

__global__ void __launch_bounds__(NTH, 2) wgrad2_kernel(const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = s32(smem);
    const int NB = a.nb;
    const int stage_bytes = (2 + NB) * BLK;
    const uint32_t s_ones = sbase + STAGES * stage_bytes;
    const uint32_t s_bar = s_ones + BLK;
    auto FULL = [&](int s) { return s_bar + 8u * s; };
    auto EMPTY = [&](int s) { return s_bar + 8u * (STAGES + s); };
    const uint32_t DONE = s_bar + 8u * (2 * STAGES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + STAGES * stage_bytes + BLK + 8 * (2 * STAGES + 1) + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long total = (long)(a.t1 - a.t0 + 1) * a.n_tiles * 2;
    const long c_begin = (long)blockIdx.x * a.chunks_per_cta;
    const long c_end = min(total, c_begin + a.chunks_per_cta);
    const bool want_bias = a.bias[0] != nullptr || a.bias[1] != nullptr;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(DONE, 1);
        fence_mbar_init();
    }
    // a block of bf16 ones (B operand of the bias column-sum UMMA)
    for (int i = tid; i < BLK / 4; i += NTH) reinterpret_cast<uint32_t *>(smem + STAGES * stage_bytes)[i] = 0x3F803F80u;
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const int nblk = (a.a_panel[1] >= 0 ? 2 : 1) + NB;
            for (long c = c_begin; c < c_end; ++c) {
                const long tt = c / 2;                       // (step, tile) index
                const int half = (int)(c & 1);
                const long step = tt / a.n_tiles, tile = tt % a.n_tiles;
                const long tbase = ((long)(a.t0 + step) * a.n_tiles + tile);
                mbar_wait(EMPTY(stage), phase ^ 1);
                mbar_expect_tx(FULL(stage), nblk * BLK);
                const uint32_t dst = sbase + stage * stage_bytes;
                tma_bulk_g2s(dst, a.A + (tbase * a.a_ppt + a.a_panel[0]) * (long)PANEL_BYTES + half * BLK, BLK, FULL(stage));
                if (a.a_panel[1] >= 0)
                    tma_bulk_g2s(dst + BLK, a.A + (tbase * a.a_ppt + a.a_panel[1]) * (long)PANEL_BYTES + half * BLK, BLK, FULL(stage));
                for (int j = 0; j < NB; ++j)
                    tma_bulk_g2s(dst + (2 + j) * BLK, a.B + (tbase * a.b_ppt + a.b_panel[j]) * (long)PANEL_BYTES + half * BLK, BLK, FULL(stage));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idN = idesc2(NB * 64, 1, 1), id16 = idesc2(16, 1, 1);
            uint32_t stage = 0, phase = 0;
            // MN-major SW128 blocks: LBO = 8 KB (next 64-column block), SBO = 1024 B (8 k-rows)
            auto desc = [](uint32_t saddr) -> uint64_t {
                return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(BLK >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            };
            bool first = true;
            for (long c = c_begin; c < c_end; ++c) {
                mbar_wait(FULL(stage), phase);
                tc_fence_after();
                const uint32_t sa = sbase + stage * stage_bytes, sb = sa + 2 * BLK;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc_mma(tmem, desc(sa + k * 2048), desc(sb + k * 2048), idN, first && k == 0 ? 0u : 1u);
                    if (want_bias) tc_mma(tmem + 240, desc(sa + k * 2048), desc(s_ones + k * 2048), id16, first && k == 0 ? 0u : 1u);
                }
                first = false;
                tc_commit(EMPTY(stage));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit(DONE);
        }
    } else {
        mbar_wait(DONE, 0);
        tc_fence_after();
        if (c_end > c_begin) {
            const int q = warp, blk = q >> 1, r = (q & 1) * 32 + lane;     // TMEM lane 32q+lane = A column = C row
            const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
            if (a.a_panel[blk] >= 0 && a.C[blk]) {
                float *dst = a.C[blk] + (long)r * a.ldc;
                for (int cc = 0; cc < NB * 64; cc += 32) {
                    uint32_t v[32];
                    tc_ld32(t_lane + cc, v);
                    tc_wait_ld();
#pragma unroll
                    for (int x = 0; x < 32; x += 4)      // 128-bit vector reductions (sm_90+): 4x fewer L2 atomic operations
                        atomicAdd(reinterpret_cast<float4 *>(dst + cc + x),
                                  make_float4(__uint_as_float(v[x]), __uint_as_float(v[x + 1]), __uint_as_float(v[x + 2]), __uint_as_float(v[x + 3])));
                }
                if (a.bias[blk]) {
                    uint32_t w[16];
                    tc_ld16(t_lane + 240, w);
                    tc_wait_ld();
                    atomicAdd(a.bias[blk] + (long)r * a.bias_stride, __uint_as_float(w[0]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
    }
}

}  // namespace w2
}  // namespace bmp

using namespace bmp;

// One contraction over the panel stash.  a_panel[1] may be -1 (single 64-column block).  N = 64 * nb <= 192
// (columns 240..255 of the accumulator hold the bias sums).
int bmp_wgrad_panels(const void *A, int a_ppt, const int a_panel[2], const void *B, int b_ppt, const int *b_panel, int nb,
                     float *const C[2], int ldc, float *const bias[2], int bias_stride,
                     int t0, int t1, int n_tiles, void *stream) {
    if (!A || !B || nb < 1 || nb > 3) { set_error("bmp_wgrad_panels: bad arguments"); return BMP_EINVAL; }
    if (t1 < t0 || n_tiles <= 0) return BMP_OK;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    w2::Args k = {};
    k.A = (const uint8_t *)A; k.B = (const uint8_t *)B; k.a_ppt = a_ppt; k.b_ppt = b_ppt;
    k.a_panel[0] = a_panel[0]; k.a_panel[1] = a_panel[1];
    for (int j = 0; j < 4; ++j) k.b_panel[j] = j < nb ? b_panel[j] : 0;
    k.nb = nb; k.C[0] = C[0]; k.C[1] = C[1]; k.ldc = ldc; k.bias[0] = bias ? bias[0] : nullptr; k.bias[1] = bias ? bias[1] : nullptr;
    k.bias_stride = bias_stride; k.t0 = t0; k.t1 = t1; k.n_tiles = n_tiles;
    const long total = (long)(t1 - t0 + 1) * n_tiles * 2;
    static int mult = 0;
    if (!mult) { const char *e = getenv("BMP_W2_CTAS"); mult = e ? atoi(e) : 1; if (mult < 1) mult = 1; }
    long ctas = (long)mult * sms;
    if (ctas > (total + 7) / 8) ctas = (total + 7) / 8;
    if (ctas < 1) ctas = 1;
    k.chunks_per_cta = (total + ctas - 1) / ctas;
    ctas = (total + k.chunks_per_cta - 1) / k.chunks_per_cta;
    const int smem = w2::STAGES * (2 + nb) * w2::BLK + w2::BLK + 256 + 1024;
    cudaFuncSetAttribute(w2::wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, w2::STAGES * 5 * w2::BLK + w2::BLK + 256 + 1024);
    w2::wgrad2_kernel<<<(unsigned)ctas, w2::NTH, smem, (cudaStream_t)stream>>>(k);
    count_launch();
    return check_launch("wgrad2_kernel");
}
