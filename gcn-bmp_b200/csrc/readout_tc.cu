// readout_tc.cu -- gated sum readout on tcgen05 (BMP_MODE_BF16), forward and backward.
// R1: models/readout/ggnn_readout.py:42-58; R2: models/ggnn_att.py:338-346 (same math as readout.cu).
//
// A tile = two padded molecules = 128 rows.  X = [h | h0] is loaded coalesced from global, converted to
// bf16 SW128 K-major panels; U = X W_i^T and V = X' W_j^T run as M=128 x N=O UMMAs into TMEM with the
// packed weight tiles streamed through a TMA ring; the epilogue applies sigmoid(u) * act(v) * mask and
// reduces over the atoms of each molecule with a 31-shuffle reduce-scatter (lane L ends up with column L).
// Backward recomputes U, V the same way, forms du / dv (fp32 to DU / DV for the weight-gradient
// contractions + bf16 operand panels), and dX = du W_i + dv W_j is a second UMMA group whose result is
// accumulated into the caller's dh / dh0 rows.
#include "tc_common.cuh"

namespace bmp {
namespace rtc {
using namespace tc;

constexpr int STAGES = 3;
// shared-memory plan (hidden / out_dim 256 is forward-only: 8 X panels, 2 weight stages of 32 KB, no transposition staging)
template <int H, int O> struct Plan {
    static constexpr bool BIG = H == 256 || O == 256;
    static constexpr int ST = BIG ? 2 : STAGES;
    static constexpr int NXP = BIG ? 8 : 4;
    static constexpr int WT = (O > H ? O : H) * 128;
    static constexpr int OFF_W = NXP * PANEL_BYTES;
    static constexpr int OFF_STG = OFF_W + ST * WT;
    static constexpr int OFF_RED = OFF_STG + (BIG ? 0 : 8 * 4096);
    static constexpr int OFF_BAR = OFF_RED + 4 * O * 4;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int MAXC = 2 * O > 2 * H ? 2 * O : 2 * H;
    static constexpr int TMEM_COLS = MAXC <= 128 ? 128 : (MAXC <= 256 ? 256 : 512);
};

struct Args {
    int mb, N, H, O, Kcat, Kj, act, act_agg;
    const float *h, *h0, *mask, *b_i, *b_j;
    const uint8_t *img;          // forward tiles [Kcat/64 of W_i][Kj/64 of W_j], then backward tiles
    float *g;                    // forward out (mb, O)
    const float *gin, *dg;       // backward in
    float *DU, *DV, *dh, *dh0;
};

// lane L of the warp ends with sum over the 32 lanes of v[L]  (31 shuffles)
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float keep = up ? v[i + off] : v[i];
            const float send = up ? v[i] : v[i + off];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <int H, int O, bool BWD>
__global__ void __launch_bounds__(NTHR, 1) readout_tc_kernel(const Args a) {
    constexpr int TILE_BYTES = O * 128;               // forward weight tile [O n][64 k]
    constexpr int BT_BYTES = H * 128;                 // backward weight tile [H n][64 k]
    using PL = Plan<H, O>;
    static_assert(!(BWD && PL::BIG), "hidden / out_dim 256 is forward-only");
    constexpr int WT = PL::WT, STAGES = PL::ST;
    constexpr int NC = O / 2;                         // U/V columns per epilogue thread
    constexpr int OFF_X = 0;                          // X panels, later du | dv panels
    constexpr int OFF_W = PL::OFF_W;
    constexpr int OFF_STG = PL::OFF_STG;
    constexpr int OFF_RED = PL::OFF_RED;              // [4 quarters][O] column partial sums
    constexpr int OFF_BAR = PL::OFF_BAR;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = s32(smem), s_x = sbase + OFF_X, s_w = sbase + OFF_W, s_bar = sbase + OFF_BAR;
    auto BAR = [&](int i) { return s_bar + 8u * i; };
    constexpr int B_FULL = 0, B_EMPTY = 4, B_XRDY = 8, B_UV = 9, B_DRDY = 10, B_DX = 11, NBAR = 12;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * NBAR + 8);
    float *red = reinterpret_cast<float *>(smem + OFF_RED);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (a.mb + 1) / 2;
    const int KPX = a.Kcat / 64, KPJ = a.Kj / 64;     // K panels of W_i / W_j
    const int n_fwd = KPX + KPJ;
    const int n_bwd = (O / 64) * ((a.Kcat / H) + (a.Kj / H));   // per H-wide output block: du part, dv part
    constexpr int TMEM_COLS = PL::TMEM_COLS;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
        mbar_init(BAR(B_XRDY), NEPI);
        mbar_init(BAR(B_UV), 1);
        mbar_init(BAR(B_DRDY), NEPI);
        mbar_init(BAR(B_DX), 1);
        fence_mbar_init();
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const uint8_t *src = a.img;
                for (int s = 0; s < n_fwd + (BWD ? n_bwd : 0); ++s) {
                    const uint32_t bytes = s < n_fwd ? TILE_BYTES : BT_BYTES;
                    mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
                    mbar_expect_tx(BAR(B_FULL + stage), bytes);
                    tma_bulk_g2s(s_w + stage * WT, src, bytes, BAR(B_FULL + stage));
                    src += bytes;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            constexpr uint32_t ID_F = idesc2(O, 0, 0), ID_B = idesc2(H, 0, 0);
            uint32_t stage = 0, phase = 0, it = 0;
            auto mma_wtile = [&](uint32_t a_addr, uint32_t dcol, bool first, uint32_t id) {
                mbar_wait(BAR(B_FULL + stage), phase);
                tc_fence_after();
                const uint32_t b_addr = s_w + stage * WT;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma(tmem + dcol, desc_kmajor(a_addr + k * 32), desc_kmajor(b_addr + k * 32), id, (first && k == 0) ? 0u : 1u);
                tc_commit(BAR(B_EMPTY + stage));
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t par = it & 1;
                mbar_wait(BAR(B_XRDY), par);
                tc_fence_after();
                for (int kp = 0; kp < KPX; ++kp) mma_wtile(s_x + kp * PANEL_BYTES, 0, kp == 0, ID_F);        // U
                for (int kp = 0; kp < KPJ; ++kp) mma_wtile(s_x + kp * PANEL_BYTES, O, kp == 0, ID_F);        // V
                tc_commit(BAR(B_UV));
                if (BWD) {
                    mbar_wait(BAR(B_DRDY), par);
                    tc_fence_after();
                    // dX block nb (H columns): du W_i[:, nb*H : (nb+1)*H]  (+ dv W_j[:, same] when W_j has that block)
                    for (int nb = 0; nb < a.Kcat / H; ++nb) {
                        for (int kp = 0; kp < O / 64; ++kp) mma_wtile(s_x + kp * PANEL_BYTES, nb * H, kp == 0, ID_B);
                        if (nb < a.Kj / H)
                            for (int kp = 0; kp < O / 64; ++kp) mma_wtile(s_x + (O / 64 + kp) * PANEL_BYTES, nb * H, false, ID_B);
                    }
                    tc_commit(BAR(B_DX));
                }
            }
        }
    } else {
        const int q = warp & 3, hf = warp >> 2;
        const int row = 32 * q + lane, colbase = hf * NC;
        const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
        const int molslot = row >> 6, atom = row & 63;
        float *stg = reinterpret_cast<float *>(smem + OFF_STG + warp * 4096);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t par = it & 1;
            const int molg = tile * 2 + molslot;
            const bool live = molg < a.mb && atom < a.N;
            const long grow = (long)molg * a.N + atom;
            auto wrow = [&](int r) -> long {
                const int tr = 32 * q + r, mg = tile * 2 + (tr >> 6), at = tr & 63;
                return (mg < a.mb && at < a.N) ? (long)mg * a.N + at : -1L;
            };
            // ---- X = [h | h0] -> bf16 panels (coalesced float4 loads, 8 in flight per thread) ----
            for (int kp0 = 0; kp0 < KPX; kp0 += 2) {          // two panels (16 float4 per thread) in flight
                float4 v[2][8];
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int kp = kp0 + p;
                    const float *src = (kp * 64 < H) ? a.h : a.h0;
                    const int c0 = (kp * 64) % H;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = u * NEPI + tid, r = idx >> 4, c4 = (idx & 15) * 4;
                        const int mg = tile * 2 + (r >> 6), at = r & 63;
                        v[p][u] = (kp < KPX && mg < a.mb && at < a.N) ? __ldg(reinterpret_cast<const float4 *>(src + ((long)mg * a.N + at) * H + c0 + c4))
                                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    if (kp0 + p >= KPX) break;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int idx = u * NEPI + tid, r = idx >> 4, c4 = (idx & 15) * 4;
                        *reinterpret_cast<uint2 *>(smem + OFF_X + (kp0 + p) * PANEL_BYTES + sw128(r, c4)) =
                            make_uint2(pack_bf16(v[p][u].x, v[p][u].y), pack_bf16(v[p][u].z, v[p][u].w));
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(BAR(B_XRDY));
            const float mk = live ? (a.mask ? a.mask[grow] : 1.f) : 0.f;
            mbar_wait(BAR(B_UV), par);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < NC; cc += 32) {
                uint32_t vu[32], vv[32];
                tc_ld32(t_lane + colbase + cc, vu);
                tc_ld32(t_lane + O + colbase + cc, vv);
                tc_wait_ld();
                float bi[32], bj[32];          // biases of this thread's columns (16-byte loads)
#pragma unroll
                for (int x = 0; x < 32; x += 4) {
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 p4 = a.b_i ? __ldg(reinterpret_cast<const float4 *>(a.b_i + colbase + cc + x)) : z4;
                    const float4 q4 = a.b_j ? __ldg(reinterpret_cast<const float4 *>(a.b_j + colbase + cc + x)) : z4;
                    bi[x] = p4.x; bi[x + 1] = p4.y; bi[x + 2] = p4.z; bi[x + 3] = p4.w;
                    bj[x] = q4.x; bj[x + 1] = q4.y; bj[x + 2] = q4.z; bj[x + 3] = q4.w;
                }
                if (!BWD) {
                    float gv[32];
#pragma unroll
                    for (int x = 0; x < 32; ++x) {
                        const float u = __uint_as_float(vu[x]) + bi[x];
                        const float v = __uint_as_float(vv[x]) + bj[x];
                        gv[x] = mk * sigmoid_fast(u) * act_fast(a.act, v);
                    }
                    const float s = warp_reduce_scatter32(gv, lane);
                    red[q * O + colbase + cc + lane] = s;
                } else {
                    float du[32], dv[32];
                    const bool mlive = molg < a.mb;
#pragma unroll
                    for (int x = 0; x < 32; ++x) {
                        const int col = colbase + cc + x;
                        const float u = __uint_as_float(vu[x]) + bi[x];
                        const float v = __uint_as_float(vv[x]) + bj[x];
                        float ds = 0.f;
                        if (mlive) {
                            const float go = __ldg(a.gin + (long)molg * O + col);
                            ds = __ldg(a.dg + (long)molg * O + col) * act_bwd(a.act_agg, go, go);
                        }
                        const float su = sigmoid_fast(u), av = act_fast(a.act, v);
                        du[x] = ds * mk * av * su * (1.f - su);
                        dv[x] = ds * mk * su * act_bwd(a.act, v, av);
                    }
                    warp_store_rows<32>(stg, du, lane, [&](int r) -> float * { const long g = wrow(r); return g >= 0 ? a.DU + g * O + colbase + cc : nullptr; });
                    warp_store_rows<32>(stg, dv, lane, [&](int r) -> float * { const long g = wrow(r); return g >= 0 ? a.DV + g * O + colbase + cc : nullptr; });
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int kk = colbase + cc + 8 * g;
                        uint4 pu = make_uint4(pack_bf16(du[8 * g], du[8 * g + 1]), pack_bf16(du[8 * g + 2], du[8 * g + 3]),
                                              pack_bf16(du[8 * g + 4], du[8 * g + 5]), pack_bf16(du[8 * g + 6], du[8 * g + 7]));
                        uint4 pv = make_uint4(pack_bf16(dv[8 * g], dv[8 * g + 1]), pack_bf16(dv[8 * g + 2], dv[8 * g + 3]),
                                              pack_bf16(dv[8 * g + 4], dv[8 * g + 5]), pack_bf16(dv[8 * g + 6], dv[8 * g + 7]));
                        *reinterpret_cast<uint4 *>(smem + OFF_X + (kk >> 6) * PANEL_BYTES + sw128(row, kk & 63)) = pu;
                        *reinterpret_cast<uint4 *>(smem + OFF_X + (O / 64 + (kk >> 6)) * PANEL_BYTES + sw128(row, kk & 63)) = pv;
                    }
                }
            }
            if (!BWD) {
                tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(NEPI));
                // combine the two row-quarters of each molecule, apply the aggregate activation
                for (int idx = tid; idx < 2 * O; idx += NEPI) {
                    const int ms = idx / O, col = idx % O, mg = tile * 2 + ms;
                    if (mg < a.mb) a.g[(long)mg * O + col] = act_fwd(a.act_agg, red[(2 * ms) * O + col] + red[(2 * ms + 1) * O + col]);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NEPI));      // red[] and the X panels are free again
            } else {
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(BAR(B_DRDY));
                mbar_wait(BAR(B_DX), par);
                tc_fence_after();
                // dh (+=) from columns [0,H), dh0 (+=) from [H,2H)
                for (int nb = 0; nb < a.Kcat / H; ++nb) {
                    float *dst = nb == 0 ? a.dh : a.dh0;
                    if (!dst) continue;
#pragma unroll
                    for (int cc = 0; cc < H / 2; cc += 16) {
                        uint32_t w[16];
                        const int col = hf * (H / 2) + cc;
                        tc_ld16(t_lane + nb * H + col, w);
                        float cur[16];
                        warp_load_rows<16>(stg, cur, lane, [&](int r) -> const float * { const long g = wrow(r); return g >= 0 ? dst + g * H + col : nullptr; });
                        tc_wait_ld();
#pragma unroll
                        for (int x = 0; x < 16; ++x) cur[x] += __uint_as_float(w[x]);
                        warp_store_rows<16>(stg, cur, lane, [&](int r) -> float * { const long g = wrow(r); return g >= 0 ? dst + g * H + col : nullptr; });
                    }
                }
                tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(NEPI));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

// packed image: forward tiles [O n][64 k]: W_i panels (Kcat/64), W_j panels (Kj/64);
// backward tiles [H n][64 k]: for nb in Kcat/H: { W_i^T block nb: O/64 panels ; if nb < Kj/H: W_j^T block nb: O/64 panels }
struct PackArgs {
    int H, O, Kcat, Kj;
    const float *W_i, *W_j;
    uint8_t *img;
};
__global__ void pack_readout_kernel(const PackArgs p) {
    const int H = p.H, O = p.O;
    const int n_fwd = p.Kcat / 64 + p.Kj / 64;
    const long fwd_elems = (long)n_fwd * O * 64;
    const int nbx = p.Kcat / H, nbj = p.Kj / H, opan = O / 64;
    const long bwd_elems = (long)(nbx + nbj) * opan * H * 64;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < fwd_elems + bwd_elems; idx += (long)gridDim.x * blockDim.x) {
        float w;
        size_t off;
        if (idx < fwd_elems) {
            const int k = idx & 63, n = (idx >> 6) % O, tile = (int)(idx / (64L * O));
            if (tile < p.Kcat / 64) w = p.W_i[(long)n * p.Kcat + tile * 64 + k];
            else w = p.W_j[(long)n * p.Kj + (tile - p.Kcat / 64) * 64 + k];
            off = (size_t)tile * O * 128 + (size_t)n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1));
        } else {
            const long j = idx - fwd_elems;
            const int k = j & 63, n = (j >> 6) % H, tile = (int)(j / (64L * H));
            // walk the block order to find (nb, which, kp)
            int t = tile, nb = 0, which = 0, kp = 0;
            for (nb = 0; nb < nbx; ++nb) {
                if (t < opan) { which = 0; kp = t; break; }
                t -= opan;
                if (nb < nbj) {
                    if (t < opan) { which = 1; kp = t; break; }
                    t -= opan;
                }
            }
            const int o = kp * 64 + k;                       // reduction index = output unit of the linear
            w = which == 0 ? p.W_i[(long)o * p.Kcat + nb * H + n] : p.W_j[(long)o * p.Kj + nb * H + n];
            off = (size_t)n_fwd * O * 128 + (size_t)tile * H * 128 + (size_t)n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1));
        }
        *reinterpret_cast<__nv_bfloat16 *>(p.img + off) = __float2bfloat16_rn(w);
    }
}

}  // namespace rtc
}  // namespace bmp

using namespace bmp;

extern "C" size_t bmp_readout_tc_workspace_bytes(int hidden, int out_dim) {
    if (hidden == 256 && out_dim == 256) return (size_t)16 * 256 * 128 + (size_t)16 * 256 * 128 + 1024;   // forward only
    if ((hidden != 64 && hidden != 128) || (out_dim != 64 && out_dim != 128)) return 0;
    return (size_t)8 * out_dim * 128 + (size_t)8 * hidden * 128 + 1024;
}

template <int H, int O>
static int launch_readout_tc(const rtc::Args &k, bool bwd, int grid, cudaStream_t st) {
    constexpr int smem = rtc::Plan<H, O>::SMEM;
    ProfScope prof(BMP_PROF_READOUT, st);
    if (bwd) {
        if constexpr (rtc::Plan<H, O>::BIG) {
            set_error("readout tcgen05 path: hidden / out_dim 256 is forward-only");
            return BMP_ESHAPE;
        } else {
            cudaFuncSetAttribute(rtc::readout_tc_kernel<H, O, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            rtc::readout_tc_kernel<H, O, true><<<grid, tc::NTHR, smem, st>>>(k);
        }
    } else {
        cudaFuncSetAttribute(rtc::readout_tc_kernel<H, O, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        rtc::readout_tc_kernel<H, O, false><<<grid, tc::NTHR, smem, st>>>(k);
    }
    count_launch();
    return check_launch("readout_tc_kernel");
}

// Shared driver for bmp_readout_forward / _backward in BMP_MODE_BF16.  Returns BMP_ESHAPE when the shape is
// outside the tensor-core kernel so the caller can use the fp32 kernel.
int bmp_readout_tc(int mb, int N, int H, int O, int variant, int act, int act_agg, const float *h, const float *h0,
                   const float *mask, const float *W_i, const float *b_i, const float *W_j, const float *b_j,
                   float *g, const float *dg, float *DU, float *DV, float *dh, float *dh0, void *ws, size_t ws_bytes,
                   bool images_ready, bool bwd, void *stream) {
    const bool big = H == 256 && O == 256;
    if ((!big && ((H != 64 && H != 128) || (O != 64 && O != 128))) || variant == BMP_READOUT_SUM || N > BMP_MAX_ATOMS) {
        set_error("readout tcgen05 path: unsupported shape H=%d O=%d variant=%d", H, O, variant);
        return BMP_ESHAPE;
    }
    if (!ws || ws_bytes < bmp_readout_tc_workspace_bytes(H, O)) { set_error("readout tcgen05 path: workspace too small"); return BMP_EINVAL; }
    if (!aligned16({b_i, b_j, h, h0})) { set_error("readout tcgen05 path: h, h0, b_i, b_j must be 16-byte aligned"); return BMP_ESHAPE; }
    cudaStream_t st = (cudaStream_t)stream;
    rtc::Args k = {};
    k.mb = mb; k.N = N; k.H = H; k.O = O; k.Kcat = h0 ? 2 * H : H; k.Kj = variant == BMP_READOUT_R2 ? H : k.Kcat;
    k.act = act; k.act_agg = act_agg; k.h = h; k.h0 = h0; k.mask = mask; k.b_i = b_i; k.b_j = b_j;
    k.g = bwd ? nullptr : g; k.gin = g; k.dg = dg; k.DU = DU; k.DV = DV; k.dh = dh; k.dh0 = dh0;
    uint8_t *img = (uint8_t *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    k.img = img;
    rtc::PackArgs p;
    p.H = H; p.O = O; p.Kcat = k.Kcat; p.Kj = k.Kj; p.W_i = W_i; p.W_j = W_j; p.img = img;
    int rc = BMP_OK;
    if (!images_ready) {
        rtc::pack_readout_kernel<<<32, 256, 0, st>>>(p);
        count_launch();
        if ((rc = check_launch("pack_readout_kernel"))) return rc;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = (mb + 1) / 2;
    const int grid = n_tiles < sms ? n_tiles : sms;
    if (big) return launch_readout_tc<256, 256>(k, bwd, grid, st);
    if (H == 64 && O == 64) return launch_readout_tc<64, 64>(k, bwd, grid, st);
    if (H == 64 && O == 128) return launch_readout_tc<64, 128>(k, bwd, grid, st);
    if (H == 128 && O == 64) return launch_readout_tc<128, 64>(k, bwd, grid, st);
    return launch_readout_tc<128, 128>(k, bwd, grid, st);
}
