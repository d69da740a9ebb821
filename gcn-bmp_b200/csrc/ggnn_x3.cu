// ggnn_x3.cu -- BMP_MODE_F32 GGNN encoder with every H x H contraction on tcgen05 at fp32-grade accuracy.
//
// Same arithmetic as ggnn.cu (models/update/ggnn_update.py:31-63, models/models/ggnn.py:72-106) and the same fp32 stash
// (Hs / Ms / Gs / RSs / Ps / dHs as include/gcnbmp.h describes them), but the dense products run on the tensor cores:
// every fp32 operand is split into a bf16 hi / lo pair (x = hi + lo up to 2^-17 relative) and each product is three
// UMMAs, hi.hi + lo.hi + hi.lo, into an fp32 TMEM accumulator (the lo.lo term is below fp32 rounding).  The step is a
// short sequence of launches over the FLAT row space (rows = mb * N atoms; the GEMMs never see molecule boundaries):
//
//   forward step t        agg_fwd      AH_e = A_e h_t (sparse rows of the dense adjacency, FFMA), deg_e = A_e 1
//                         rowgemm3     m   = [AH_0 .. AH_3] [W_0 .. W_3]^T + sum_e deg_e b_e           -> Ms[t]
//                         rowgemm3     r,z = sigma([h | m] [W_r + U_r | ..]^T + b), r*h                -> Gs[t], RSs[t]
//                         rowgemm3     hb  = tanh([h | m | r*h] [W | U]^T + b), h' = z hb + (1 - z) h  -> Gs[t], Hs[t+1]
//   backward step t       gate_bwd     delta_z, delta_h, g (1 - z)            (pointwise)
//                         rowgemm3     q = delta_h U -> delta_r, ds += q r
//                         rowgemm3     dh_x = [dz | dh | dr] [W_z + U_z | W | W_r + U_r]_h (+ ds) ; dm = [..] [..]_m
//                         agg_bwd      P_e = A_e^T dm                                                  -> Ps[t]
//                         rowgemm3     dHs[t] += dh_x + sum_e P_e W_e
// The parameter gradients stay where they were: bmp_ggnn_backward's contractions over the stash (bmp_wgrad_tc3).
//
// rowgemm3: one persistent CTA per SM; a work item = (128-row tile, job), jobs of one launch interleaved so the CTAs that
// share an A tile run at the same time (second reader hits L2).  Warp roles: 8 epilogue warps (TMEM lane quarter x column half),
// 6 converter warps (fp32 rows -> hi / lo SW128 K-major panels; two k-tiles of cp.async copies in flight per CTA in
// thread-private shared-memory slots, later k-tiles prefetched towards L2), one producer lane (packed hi / lo weight k-tiles, cp.async.bulk + mbarrier), one MMA lane.  Two 128-column
// TMEM accumulators alternate between consecutive items, so an item's epilogue overlaps the next item's MMAs.
#include <cstring>
#include "tc_common.cuh"

namespace bmp {
namespace x3 {
using namespace tc;

constexpr int MAXB = 4, MAXJ = 4;
constexpr int STAGES = 2;
constexpr int STAGE_BYTES = 4 * PANEL_BYTES;      // [A_hi | A_lo | W_hi | W_lo]
constexpr int F32_DEPTH = 2;                      // fp32 A k-tiles in flight per CTA (cp.async into thread-private slots)
constexpr int WSLOT = 2 * PANEL_BYTES;            // one packed weight k-tile: hi (n x 64 bf16, SW128) then lo
constexpr int NEPI_W = 8, NCONV_W = 6;
constexpr int NT = 32 * (NEPI_W + NCONV_W + 2);   // 512 threads: 128 registers each
constexpr int NCONV = 32 * NCONV_W;
constexpr int NU = (2048 + NCONV - 1) / NCONV;    // float4 of a 128 x 64 fp32 k-tile per converter thread (last round partial)
constexpr int OFF_F32 = STAGES * STAGE_BYTES;     // [depth][NU][NCONV threads] float4
constexpr int OFF_STG = OFF_F32 + F32_DEPTH * NU * NCONV * 16;     // epilogue transposition staging: 2 KB per warp
constexpr int OFF_BAR = OFF_STG + NEPI_W * 2048;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;

enum { EPI_MSG = 0, EPI_R, EPI_Z, EPI_HB, EPI_Q, EPI_DHX, EPI_DM, EPI_DHMSG };

struct Job {
    const float *A[MAXB];            // K-block b: (rows, 64 * kt[b]) fp32, leading dimension lda[b]
    int lda[MAXB], kt[MAXB], nblk;
    const uint8_t *wimg;             // packed weight k-tiles of this job (and column chunk) in consumption order
    int epi;
    const float *bias, *bias2;       // per output column (pointers already at column n0), either may be NULL
    const float *in0, *in1, *in2;    // epilogue inputs, row-major, pointers at (row 0, column n0)
    int li0, li1, li2;
    float *out0, *out1, *out2;
    int lo0, lo1, lo2;
    const float *deg, *msg_b;        // EPI_MSG: deg (rows, 4); msg_b + n0 * 4 (the reference's b[c*E+e])
};
struct Args {
    Job job[MAXJ];
    int njobs, NC;                   // NC = columns per job: 64 or 128
    long rows;
    long long *dbg;                  // optional wait-cycle counters of CTA 0 (tools/x3_check.py --waits)
};

// ------------------------------------------------------------------------------------------------ rowgemm3
__device__ __forceinline__ float sigmoid_x3(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// 1 - 2 / (1 + e^{2x}): absolute error at the fp32 rounding level of the result's range
__device__ __forceinline__ float tanh_x3(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
// wait with a sleep between polls: a waiting warp leaves the issue slots to the warps that have work
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "nanosleep.u32 %2;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <bool DBG>
__global__ void __launch_bounds__(NT, 1) rowgemm3_kernel(const __grid_constant__ Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = s32(smem);
    const uint32_t s_bar = sbase + OFF_BAR;
    auto FULL = [&](int s) { return s_bar + 8u * s; };
    auto EMPTY = [&](int s) { return s_bar + 8u * (STAGES + s); };
    auto ACCF = [&](int b) { return s_bar + 8u * (2 * STAGES + b); };
    auto ACCE = [&](int b) { return s_bar + 8u * (2 * STAGES + 2 + b); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * (2 * STAGES + 4) + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NC = a.NC, njobs = a.njobs;
    const long nrows = a.rows;
    const long ntiles = (nrows + 127) / 128, nitems = ntiles * njobs;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), NCONV_W + 1); mbar_init(EMPTY(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(ACCF(b), 1); mbar_init(ACCE(b), NEPI_W); }
        fence_mbar_init();
    }
    if (warp == NEPI_W + NCONV_W + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == NEPI_W + NCONV_W) {
        // ===================== producer: packed weight k-tiles
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t wb = (uint32_t)NC * 128u;
            for (long w = blockIdx.x; w < nitems; w += gridDim.x) {
                const Job &J = a.job[(int)(w % njobs)];
                int nk = 0;
                for (int b = 0; b < J.nblk; ++b) nk += J.kt[b];
                for (int k = 0; k < nk; ++k) {
                    mbar_wait_sleep(EMPTY(stage), phase ^ 1, 100);
                    mbar_expect_tx(FULL(stage), 2 * wb);
                    const uint32_t dst = sbase + stage * STAGE_BYTES + 2 * PANEL_BYTES;
                    tma_bulk_g2s(dst, J.wimg + (size_t)k * WSLOT, wb, FULL(stage));
                    tma_bulk_g2s(dst + PANEL_BYTES, J.wimg + (size_t)k * WSLOT + PANEL_BYTES, wb, FULL(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == NEPI_W + NCONV_W + 1) {
        // ===================== MMA issuer
        if (lane == 0) {
            const uint32_t ID = idesc(NC, 0);
            uint32_t stage = 0, phase = 0, it = 0;
            const bool dbg = DBG && blockIdx.x == 0;
            long long t_all = dbg ? clock64() : 0, w_full = 0, w_acce = 0, t0 = 0;
            for (long w = blockIdx.x; w < nitems; w += gridDim.x, ++it) {
                const Job &J = a.job[(int)(w % njobs)];
                int nk = 0;
                for (int b = 0; b < J.nblk; ++b) nk += J.kt[b];
                const uint32_t buf = it & 1, use = it >> 1;
                if (dbg) t0 = clock64();
                mbar_wait_sleep(ACCE(buf), (use & 1) ^ 1, 100);           // the epilogue has drained this accumulator's previous item
                if (dbg) w_acce += clock64() - t0;
                tc_fence_after();
                const uint32_t d = tmem + buf * 128;
                for (int k = 0; k < nk; ++k) {
                    if (dbg) t0 = clock64();
                    mbar_wait_sleep(FULL(stage), phase, 32);
                    if (dbg) w_full += clock64() - t0;
                    tc_fence_after();
                    const uint32_t sa = sbase + stage * STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ah = desc_kmajor(sa + ks * 32), al = desc_kmajor(sa + PANEL_BYTES + ks * 32);
                        const uint64_t wh = desc_kmajor(sa + 2 * PANEL_BYTES + ks * 32), wl = desc_kmajor(sa + 3 * PANEL_BYTES + ks * 32);
                        tc_mma(d, ah, wh, ID, (k | ks) ? 1u : 0u);      // hi . hi
                        tc_mma(d, al, wh, ID, 1u);                      // lo . hi
                        tc_mma(d, ah, wl, ID, 1u);                      // hi . lo
                    }
                    tc_commit(EMPTY(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(ACCF(buf));
            }
            if (dbg) { a.dbg[0] = clock64() - t_all; a.dbg[1] = w_full; a.dbg[2] = w_acce; }
        }
    } else if (warp >= NEPI_W) {
        // ===================== converters: fp32 rows -> bf16 hi / lo K-major panels
        // Thread ct owns the 16-byte piece p = ct & 15 of the rows r0 + 12 u (r0 = ct >> 4, u = 0..10; the last round
        // only for ct < 128): global source, private slot and swizzled panel offset are a base plus a constant per round.
        const int ct = tid - 32 * NEPI_W;
        const int r0 = ct >> 4, p = ct & 15;
        const bool last_round = ct < 2048 - (NU - 1) * NCONV;
        const uint32_t sw_a = (uint32_t)r0 * 128u + ((((uint32_t)(p >> 1)) ^ ((uint32_t)r0 & 7u)) << 4) + (uint32_t)(p & 1) * 8u;
        const uint32_t sw_b = (uint32_t)r0 * 128u + ((((uint32_t)(p >> 1)) ^ ((uint32_t)(r0 + 4) & 7u)) << 4) + (uint32_t)(p & 1) * 8u;
        const uint32_t f32_mine = sbase + OFF_F32 + (uint32_t)ct * 16u;
        struct Cur { long w; int b, k; };
        auto valid = [&](const Cur &c) { return c.w < nitems; };
        auto advance = [&](Cur c) {
            const Job &J = a.job[(int)(c.w % njobs)];
            if (++c.k == J.kt[c.b]) { c.k = 0; if (++c.b == J.nblk) { c.b = 0; c.w += gridDim.x; } }
            return c;
        };
        // issue this thread's asynchronous 16-byte copies of a k-tile into its private slots (rows beyond the end: zero-fill)
        auto issue = [&](const Cur &c, int slot) {
            const Job &J = a.job[(int)(c.w % njobs)];
            const long row0 = (c.w / njobs) * 128;
            const long lda = J.lda[c.b];
            const float *src = J.A[c.b] + (row0 + r0) * lda + c.k * 64 + p * 4;
            const long step = 12 * lda;
            const uint32_t dst = f32_mine + (uint32_t)slot * (NU * NCONV * 16u);
            if (row0 + 128 <= nrows) {
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    if (u < NU - 1 || last_round)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)u * (NCONV * 16u)), "l"(src) : "memory");
                    src += step;
                }
            } else {
                const int left = (int)(nrows - row0);
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const bool live = r0 + 12 * u < left;
                    if (u < NU - 1 || last_round)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)u * (NCONV * 16u)),
                                     "l"(live ? src : J.A[c.b]), "r"(live ? 16 : 0) : "memory");
                    src += step;
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // pull a k-tile (and, at the first k-tile of an item, the item's epilogue inputs) towards L2: thread ct < 128 = row ct
        auto prefetch = [&](const Cur &c) {
            if (ct >= 128) return;
            const Job &J = a.job[(int)(c.w % njobs)];
            const long row = (c.w / njobs) * 128 + ct;
            if (row >= nrows) return;
            const float *src = J.A[c.b] + row * J.lda[c.b] + c.k * 64;
            prefetch_l2(src);
            prefetch_l2(src + 32);
            if (c.b == 0 && c.k == 0) {
                for (int c0 = 0; c0 < NC; c0 += 32) {
                    if (J.in0) prefetch_l2(J.in0 + row * J.li0 + c0);
                    if (J.in1) prefetch_l2(J.in1 + row * J.li1 + c0);
                    if (J.in2) prefetch_l2(J.in2 + row * J.li2 + c0);
                }
            }
        };
        constexpr int PF = 4;
        uint32_t stage = 0, phase = 0;
        Cur cur{(long)blockIdx.x, 0, 0};
        Cur pf = cur;
        for (int i = 0; i < PF && valid(pf); ++i) {
            prefetch(pf);
            pf = advance(pf);
        }
        const bool dbg = DBG && blockIdx.x == 0 && ct == 0;
        long long w_empty = 0, w_conv = 0, t0 = 0;
        // F32_DEPTH k-tiles in flight: one commit group per k-tile (empty groups past the end keep the count uniform)
        Cur ahead = cur;
        for (int i = 0; i < F32_DEPTH; ++i) {
            if (valid(ahead)) { issue(ahead, i); ahead = advance(ahead); }
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        int slot = 0;
        while (valid(cur)) {
            if (valid(pf)) {
                prefetch(pf);
                pf = advance(pf);
            }
            asm volatile("cp.async.wait_group %0;" ::"n"(F32_DEPTH - 1) : "memory");
            if (dbg) t0 = clock64();
            mbar_wait_sleep(EMPTY(stage), phase ^ 1, 100);
            if (dbg) { w_empty += clock64() - t0; t0 = clock64(); }
            const uint32_t hi = sbase + stage * STAGE_BYTES, mine = f32_mine + (uint32_t)slot * (NU * NCONV * 16u);
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                if (u == NU - 1 && !last_round) break;
                float4 x;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(mine + (uint32_t)u * (NCONV * 16u)));
                const uint32_t off = hi + ((u & 1) ? sw_b : sw_a) + (uint32_t)u * (12u * 128u);
                const uint32_t hx = pack_bf16(x.x, x.y), hy = pack_bf16(x.z, x.w);
                const uint32_t lx = pack_bf16(x.x - __uint_as_float(hx << 16), x.y - __uint_as_float(hx & 0xFFFF0000u));
                const uint32_t ly = pack_bf16(x.z - __uint_as_float(hy << 16), x.w - __uint_as_float(hy & 0xFFFF0000u));
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(off), "r"(hx), "r"(hy) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(off + PANEL_BYTES), "r"(lx), "r"(ly) : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(FULL(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            // the slot just read is free: refill it with the k-tile F32_DEPTH ahead
            if (valid(ahead)) { issue(ahead, slot); ahead = advance(ahead); }
            else asm volatile("cp.async.commit_group;" ::: "memory");
            if (++slot == F32_DEPTH) slot = 0;
            cur = advance(cur);
            if (dbg) w_conv += clock64() - t0;
        }
        if (dbg) { a.dbg[3] = w_empty; a.dbg[4] = w_conv; }
    } else {
        // ===================== epilogue: TMEM -> pointwise -> global (coalesced through a per-warp transposition block)
        // warp = TMEM lane quarter (warp & 3) x column half (warp >> 2); 16-column chunks
        float *stg = reinterpret_cast<float *>(smem + OFF_STG + warp * 2048);
        const int q = warp & 3, half = warp >> 2;
        const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
        const int nch = NC / 32;                               // chunks of this warp
        uint32_t it = 0;
        const bool dbg = DBG && blockIdx.x == 0 && tid == 0;
        long long w_accf = 0, w_epi = 0, t0 = 0;
        for (long w = blockIdx.x; w < nitems; w += gridDim.x, ++it) {
            const Job &J = a.job[(int)(w % njobs)];
            const int epi = J.epi;
            if (dbg) t0 = clock64();
            const long wrow0 = (w / njobs) * 128 + 32 * q;           // first row of this warp
            const long row = wrow0 + lane;
            const int nlive = nrows - wrow0 >= 32 ? 32 : (int)(nrows - wrow0);   // may be <= 0
            const uint32_t buf = it & 1, use = it >> 1;
            mbar_wait_sleep(ACCF(buf), use & 1, 100);
            if (dbg) { w_accf += clock64() - t0; t0 = clock64(); }
            tc_fence_after();
            for (int cc = 0; cc < nch; ++cc) {
                const int c0 = half * (NC / 2) + cc * 16;
                auto rp = [nlive, wrow0, c0](const float *p, int ld) {
                    const float *base = p + wrow0 * ld + c0;
                    return [=](int r) -> const float * { return r < nlive ? base + r * ld : nullptr; };
                };
                auto wp = [nlive, wrow0, c0](float *p, int ld) {
                    float *base = p + wrow0 * ld + c0;
                    return [=](int r) -> float * { return r < nlive ? base + r * ld : nullptr; };
                };
                // the global inputs of the chunk are requested before the accumulator is read
                float4 x0[4], x1[4];
                const bool u0 = epi == EPI_R || epi == EPI_HB || epi == EPI_Q || epi == EPI_DHMSG || (epi == EPI_DHX && J.in0);
                const bool u1 = (epi == EPI_HB && J.in1) || epi == EPI_Q || epi == EPI_DHMSG;
                if (u0) warp_ldg_rows<16>(x0, lane, rp(J.in0, J.li0));
                if (u1) warp_ldg_rows<16>(x1, lane, rp(J.in1, J.li1));
                uint32_t vr[16];
                tc_ld16(t_lane + buf * 128 + c0, vr);
                tc_wait_ld();
                if (cc == nch - 1) {                                 // every TMEM read of this item is done
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ACCE(buf));
                }
                float f[16], i0[16], i1[16];
#pragma unroll
                for (int x = 0; x < 16; ++x) f[x] = __uint_as_float(vr[x]);
                if (u0) warp_transpose_in<16>(stg, x0, i0, lane);
                if (u1) warp_transpose_in<16>(stg, x1, i1, lane);
                if (epi == EPI_R || epi == EPI_Z || epi == EPI_HB) {
#pragma unroll
                    for (int x = 0; x < 16; x += 4) {
                        if (J.bias) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(J.bias + c0 + x));
                            f[x] += b4.x; f[x + 1] += b4.y; f[x + 2] += b4.z; f[x + 3] += b4.w;
                        }
                        if (J.bias2) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(J.bias2 + c0 + x));
                            f[x] += b4.x; f[x + 1] += b4.y; f[x + 2] += b4.z; f[x + 3] += b4.w;
                        }
                    }
                }
                switch (epi) {
                    case EPI_MSG: {
                        const float4 d = row < nrows ? __ldg(reinterpret_cast<const float4 *>(J.deg + row * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int x = 0; x < 16; ++x) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(J.msg_b + (c0 + x) * 4));
                            f[x] += b4.x * d.x + b4.y * d.y + b4.z * d.z + b4.w * d.w;
                        }
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        break;
                    }
                    case EPI_R: {
#pragma unroll
                        for (int x = 0; x < 16; ++x) { f[x] = sigmoid_x3(f[x]); i0[x] *= f[x]; }
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        warp_store_rows<16>(stg, i0, lane, wp(J.out1, J.lo1));
                        break;
                    }
                    case EPI_Z: {
#pragma unroll
                        for (int x = 0; x < 16; ++x) f[x] = sigmoid_x3(f[x]);
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        if (J.out1) {           // stateless step: zero r slot and r*state (keeps the merged wgrad contractions exact)
#pragma unroll
                            for (int x = 0; x < 16; ++x) f[x] = 0.f;
                            warp_store_rows<16>(stg, f, lane, wp(J.out1, J.lo1));
                            warp_store_rows<16>(stg, f, lane, wp(J.out2, J.lo2));
                        }
                        break;
                    }
                    case EPI_HB: {              // i0 = z, i1 = state
#pragma unroll
                        for (int x = 0; x < 16; ++x) f[x] = tanh_x3(f[x]);
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        if (u1) {
#pragma unroll
                            for (int x = 0; x < 16; ++x) f[x] = i0[x] * f[x] + (1.f - i0[x]) * i1[x];
                        } else {
#pragma unroll
                            for (int x = 0; x < 16; ++x) f[x] = i0[x] * f[x];
                        }
                        warp_store_rows<16>(stg, f, lane, wp(J.out1, J.lo1));
                        if (J.out2) warp_store_rows<16>(stg, f, lane, wp(J.out2, J.lo2));
                        break;
                    }
                    case EPI_Q: {               // i0 = state, i1 = r, in2 = ds
                        float d[16];
                        warp_load_rows<16>(stg, d, lane, rp(J.in2, J.li2));
#pragma unroll
                        for (int x = 0; x < 16; ++x) {
                            d[x] += f[x] * i1[x];
                            f[x] = f[x] * i0[x] * i1[x] * (1.f - i1[x]);
                        }
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        warp_store_rows<16>(stg, d, lane, wp(J.out1, J.lo1));
                        break;
                    }
                    case EPI_DHX: {
                        if (u0) {
#pragma unroll
                            for (int x = 0; x < 16; ++x) f[x] += i0[x];
                        }
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        break;
                    }
                    case EPI_DM:
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        break;
                    default: {   // EPI_DHMSG
#pragma unroll
                        for (int x = 0; x < 16; ++x) f[x] += i0[x] + i1[x];
                        warp_store_rows<16>(stg, f, lane, wp(J.out0, J.lo0));
                        break;
                    }
                }
            }
            if (dbg) w_epi += clock64() - t0;
        }
        if (dbg) { a.dbg[5] = w_accf; a.dbg[6] = w_epi; a.dbg[7] = it; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NEPI_W + NCONV_W + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    }
}

// ------------------------------------------------------------------------------------------------ weight images
// Image of one step's parameters: the k-tiles of the eight jobs in consumption order, per 128-column output chunk.
//   job        K-blocks (each H/64 k-tiles)                           B element (n = output column, k = K index)
//   MSG        AH_0 .. AH_3                                           msg_W[(n*4 + e)*H + k]
//   R          h, m                                                   W_r[n][k] + U_r[n][k] ; W_r[n][H + k]
//   Z          h, m                                                   W_z[n][k] (+ U_z[n][k] stateful) ; W_z[n][H + k]
//   HB         h, m, r*h                                              W[n][k] ; W[n][H + k] ; U[n][k]
//   Q          delta_h                                                U[k][n]
//   DHX        delta_z, delta_h, delta_r                              W_z[k][n] (+ U_z[k][n]) ; W[k][n] ; W_r[k][n] + U_r[k][n]
//   DM         delta_z, delta_h, delta_r                              W_z[k][H + n] ; W[k][H + n] ; W_r[k][H + n]
//   DHMSG      P_0 .. P_3                                             msg_W[(k*4 + e)*H + n]
__host__ __device__ inline int job_blocks(int j) { return (j == EPI_MSG || j == EPI_DHMSG) ? 4 : (j == EPI_Q ? 1 : (j == EPI_R || j == EPI_Z) ? 2 : 3); }
__host__ __device__ inline int job_tile0(int j, int H) {      // first k-tile of job j (chunk 0) inside an image
    const int kb = H / 64, hc = (H + 127) / 128;
    int t = 0;
    for (int i = 0; i < j; ++i) t += job_blocks(i) * kb * hc;
    return t;
}
__host__ __device__ inline int image_tiles(int H) { return job_tile0(8, H); }

struct PackArgs {
    const float *msg_W;
    bmp_gru_t g;
    int H, stateful;
    uint8_t *img;
};

__global__ void __launch_bounds__(256) pack_x3_kernel(const PackArgs a) {
    const int H = a.H, kb = H / 64, hc = (H + 127) / 128, NC = H < 128 ? H : 128;
    int tile = blockIdx.x, j = 0;
    while (j < 7 && tile >= job_tile0(j + 1, H)) ++j;
    const int local = tile - job_tile0(j, H), per = job_blocks(j) * kb;
    const int nc = local / per, kk = local % per, b = kk / kb, kbase = (kk % kb) * 64;
    (void)hc;
    const bool st = a.stateful != 0;
    auto val = [&](int n, int k) -> float {       // n: global output column, k: column inside K-block b
        const bmp_gru_t &g = a.g;
        const long H2 = 2L * H;
        switch (j) {
            case EPI_MSG: return a.msg_W[((long)n * 4 + b) * H + k];
            case EPI_R: return b == 0 ? (st ? g.W_r[n * H2 + k] + g.U_r[(long)n * H + k] : 0.f) : (st ? g.W_r[n * H2 + H + k] : 0.f);
            case EPI_Z: return b == 0 ? g.W_z[n * H2 + k] + (st ? g.U_z[(long)n * H + k] : 0.f) : g.W_z[n * H2 + H + k];
            case EPI_HB: return b == 0 ? g.W[n * H2 + k] : (b == 1 ? g.W[n * H2 + H + k] : (st ? g.U[(long)n * H + k] : 0.f));
            case EPI_Q: return st ? g.U[(long)k * H + n] : 0.f;
            case EPI_DHX:
                return b == 0 ? g.W_z[k * H2 + n] + (st ? g.U_z[(long)k * H + n] : 0.f)
                              : (b == 1 ? g.W[k * H2 + n] : (st ? g.W_r[k * H2 + n] + g.U_r[(long)k * H + n] : 0.f));
            case EPI_DM: return b == 0 ? g.W_z[k * H2 + H + n] : (b == 1 ? g.W[k * H2 + H + n] : (st ? g.W_r[k * H2 + H + n] : 0.f));
            default: return a.msg_W[((long)k * 4 + b) * H + n];
        }
    };
    uint8_t *dst = a.img + (size_t)tile * WSLOT;
    for (int idx = threadIdx.x; idx < NC * 8; idx += 256) {
        const int n = idx % NC, c8 = idx / NC;        // n fastest: the transposed jobs read consecutive addresses
        float v[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) v[x] = val(nc * 128 + n, kbase + c8 * 8 + x);
        uint32_t h[4], l[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            h[x] = pack_bf16(v[2 * x], v[2 * x + 1]);
            l[x] = pack_bf16(v[2 * x] - __uint_as_float(h[x] << 16), v[2 * x + 1] - __uint_as_float(h[x] & 0xFFFF0000u));
        }
        const uint32_t off = sw128(n, c8 * 8);
        *reinterpret_cast<uint4 *>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(dst + PANEL_BYTES + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ------------------------------------------------------------------------------------------------ adjacency products
// The adjacency is the reference's dense fp32 (mb,E,N,N) array; its rows are scanned for non-zeros (a warp ballot) and only
// those neighbours are accumulated, in ascending order -- the same sum as the dense product, zeros skipped.
// AH[row][e*H + c] = sum_j A_e[i][j] h[j][c] ; deg[row][e] = sum_j A_e[i][j]        (one CTA per molecule)
__global__ void __launch_bounds__(256) agg_fwd_kernel(const float *__restrict__ adj, const float *__restrict__ h, float *__restrict__ AH,
                                                      float *__restrict__ deg, int mb, int N, int H) {
    extern __shared__ __align__(16) float sh[];                  // [N][H]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, HV = H / 4;
    for (int mol = blockIdx.x; mol < mb; mol += gridDim.x) {
        __syncthreads();
        const float4 *src = reinterpret_cast<const float4 *>(h + (long)mol * N * H);
        for (int idx = tid; idx < N * HV; idx += 256) reinterpret_cast<float4 *>(sh)[idx] = __ldg(src + idx);
        __syncthreads();
        for (int idx = warp; idx < 4 * N; idx += 8) {
            const int e = idx / N, i = idx - e * N;
            const float *arow = adj + (((long)mol * 4 + e) * N + i) * N;
            const float a0 = lane < N ? __ldg(arow + lane) : 0.f, a1 = lane + 32 < N ? __ldg(arow + lane + 32) : 0.f;
            uint32_t m0 = __ballot_sync(0xffffffffu, a0 != 0.f), m1 = __ballot_sync(0xffffffffu, a1 != 0.f);
            float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
            auto fma4 = [](float4 &acc, float s, const float4 &x) {
                acc.x = fmaf(s, x.x, acc.x); acc.y = fmaf(s, x.y, acc.y); acc.z = fmaf(s, x.z, acc.z); acc.w = fmaf(s, x.w, acc.w);
            };
            while (m0) {
                const int j = __ffs(m0) - 1;
                m0 &= m0 - 1;
                const float s = __shfl_sync(0xffffffffu, a0, j);
                const float4 *hr = reinterpret_cast<const float4 *>(sh + j * H);
                if (lane < HV) fma4(acc0, s, hr[lane]);
                if (lane + 32 < HV) fma4(acc1, s, hr[lane + 32]);
            }
            while (m1) {
                const int j = __ffs(m1) - 1;
                m1 &= m1 - 1;
                const float s = __shfl_sync(0xffffffffu, a1, j);
                const float4 *hr = reinterpret_cast<const float4 *>(sh + (j + 32) * H);
                if (lane < HV) fma4(acc0, s, hr[lane]);
                if (lane + 32 < HV) fma4(acc1, s, hr[lane + 32]);
            }
            float4 *out = reinterpret_cast<float4 *>(AH + ((long)mol * N + i) * 4 * H + (long)e * H);
            if (lane < HV) out[lane] = acc0;
            if (lane + 32 < HV) out[lane + 32] = acc1;
            if (deg) {
                float d = a0 + a1;
                for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (lane == 0) deg[((long)mol * N + i) * 4 + e] = d;
            }
        }
    }
}

// P[row j][e*H + c] = sum_i A_e[i][j] dm[i][c]      (one CTA per molecule; the bond type's tile is transposed through smem)
__global__ void __launch_bounds__(256) agg_bwd_kernel(const float *__restrict__ adj, const float *__restrict__ dm, float *__restrict__ P,
                                                      int mb, int N, int H) {
    extern __shared__ __align__(16) float sh[];                  // [N][H] dm, then [4][N][N + 1] adjacency
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, HV = H / 4, LD = N + 1;
    float *At = sh + N * H;
    for (int mol = blockIdx.x; mol < mb; mol += gridDim.x) {
        __syncthreads();
        const float4 *src = reinterpret_cast<const float4 *>(dm + (long)mol * N * H);
        for (int idx = tid; idx < N * HV; idx += 256) reinterpret_cast<float4 *>(sh)[idx] = __ldg(src + idx);
        const float *am = adj + (long)mol * 4 * N * N;
        for (int idx = tid; idx < 4 * N * N; idx += 256) {
            const int ei = idx / N, j = idx - ei * N;
            At[ei * LD + j] = __ldg(am + idx);
        }
        __syncthreads();
        for (int idx = warp; idx < 4 * N; idx += 8) {
            const int e = idx / N, j = idx - e * N;
            const float *col = At + (long)e * N * LD + j;
            const float a0 = lane < N ? col[lane * LD] : 0.f, a1 = lane + 32 < N ? col[(lane + 32) * LD] : 0.f;
            uint32_t m0 = __ballot_sync(0xffffffffu, a0 != 0.f), m1 = __ballot_sync(0xffffffffu, a1 != 0.f);
            float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
            auto fma4 = [](float4 &acc, float s, const float4 &x) {
                acc.x = fmaf(s, x.x, acc.x); acc.y = fmaf(s, x.y, acc.y); acc.z = fmaf(s, x.z, acc.z); acc.w = fmaf(s, x.w, acc.w);
            };
            while (m0) {
                const int i = __ffs(m0) - 1;
                m0 &= m0 - 1;
                const float s = __shfl_sync(0xffffffffu, a0, i);
                const float4 *hr = reinterpret_cast<const float4 *>(sh + i * H);
                if (lane < HV) fma4(acc0, s, hr[lane]);
                if (lane + 32 < HV) fma4(acc1, s, hr[lane + 32]);
            }
            while (m1) {
                const int i = __ffs(m1) - 1;
                m1 &= m1 - 1;
                const float s = __shfl_sync(0xffffffffu, a1, i);
                const float4 *hr = reinterpret_cast<const float4 *>(sh + (i + 32) * H);
                if (lane < HV) fma4(acc0, s, hr[lane]);
                if (lane + 32 < HV) fma4(acc1, s, hr[lane + 32]);
            }
            float4 *out = reinterpret_cast<float4 *>(P + ((long)mol * N + j) * 4 * H + (long)e * H);
            if (lane < HV) out[lane] = acc0;
            if (lane + 32 < HV) out[lane + 32] = acc1;
        }
    }
}

// gate derivatives of one step: g = dL/dh_{t+1};  delta_h = g z (1 - hb^2), delta_z = g (hb - s) z (1 - z) over the z | hb
// slots of Gs[t];  ds = g (1 - z) (stateful steps)
__global__ void __launch_bounds__(256) gate_bwd_kernel(const float *__restrict__ g, float *__restrict__ Gt, const float *__restrict__ s,
                                                       float *__restrict__ ds, long rows, int H) {
    const int HV = H / 4;
    const long n = rows * HV;
    for (long idx = (long)blockIdx.x * 256 + threadIdx.x; idx < n; idx += (long)gridDim.x * 256) {
        const long row = idx / HV;
        const int c = (int)(idx - row * HV) * 4;
        const float4 gv = __ldg(reinterpret_cast<const float4 *>(g + row * H + c));
        float4 *zp = reinterpret_cast<float4 *>(Gt + row * 3 * H + H + c), *hp = reinterpret_cast<float4 *>(Gt + row * 3 * H + 2 * H + c);
        const float4 z = *zp, hb = *hp;
        const float4 sv = s ? __ldg(reinterpret_cast<const float4 *>(s + row * H + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 dz, dh, d;
#define BMP_X3_GATE(f)                                   \
        dz.f = gv.f * (hb.f - sv.f) * z.f * (1.f - z.f); \
        dh.f = gv.f * z.f * (1.f - hb.f * hb.f);         \
        d.f = gv.f * (1.f - z.f);
        BMP_X3_GATE(x) BMP_X3_GATE(y) BMP_X3_GATE(z) BMP_X3_GATE(w)
#undef BMP_X3_GATE
        *zp = dz;
        *hp = dh;
        if (s) *reinterpret_cast<float4 *>(ds + row * H + c) = d;
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

static bool shape_ok(int N, int H, int E) { return (H == 64 || H == 128 || H == 256) && E == 4 && N > 0 && N <= BMP_MAX_ATOMS; }

struct Layout {
    size_t img_bytes, tmp_off, deg_off, mini_off, total;
    Layout(long rows, int H, int T, bool inference) {
        img_bytes = (size_t)image_tiles(H) * WSLOT;
        tmp_off = (size_t)T * img_bytes;
        deg_off = tmp_off + (size_t)rows * 4 * H * sizeof(float);
        mini_off = deg_off + (((size_t)rows * 4 * sizeof(float) + 1023) & ~(size_t)1023);
        total = mini_off + (inference ? (size_t)rows * 7 * H * sizeof(float) : 0) + 1024;
    }
};

static long long *g_dbg = nullptr;
static int g_dbg_slot = 0;

static int launch_gemm(Args &g, cudaStream_t st) {
    g.dbg = g_dbg ? g_dbg + 8 * (g_dbg_slot++) : nullptr;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(rowgemm3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        cudaFuncSetAttribute(rowgemm3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        attr = true;
    }
    const long items = ((g.rows + 127) / 128) * g.njobs;
    const int grid = (int)(items < sm_count() ? items : sm_count());
    if (g.dbg) rowgemm3_kernel<true><<<grid, NT, SMEM_BYTES, st>>>(g);
    else rowgemm3_kernel<false><<<grid, NT, SMEM_BYTES, st>>>(g);
    count_launch();
    return check_launch("rowgemm3_kernel");
}

// step -> image index (steps that share every parameter pointer and the stateful flag share an image)
template <class S>
static void image_plan(const S *a, int *img_of) {
    for (int t = 0; t < a->n_steps; ++t) {
        img_of[t] = t;
        for (int u = 0; u < t; ++u)
            if (a->msg_W[u] == a->msg_W[t] && same_gru(a->gru[u], a->gru[t]) && (a->stateful[u] != 0) == (a->stateful[t] != 0)) {
                img_of[t] = img_of[u];
                break;
            }
    }
}

template <class S>
static int pack_images(const S *a, uint8_t *ws, const Layout &L, const int *img_of, cudaStream_t st) {
    for (int t = 0; t < a->n_steps; ++t) {
        if (img_of[t] != t) continue;
        PackArgs p;
        p.msg_W = a->msg_W[t];
        p.g = a->gru[t];
        p.H = a->hidden;
        p.stateful = a->stateful[t];
        p.img = ws + (size_t)t * L.img_bytes;
        pack_x3_kernel<<<image_tiles(a->hidden), 256, 0, st>>>(p);
        count_launch();
    }
    return check_launch("pack_x3_kernel");
}

static const uint8_t *job_img(const uint8_t *img, int j, int nc, int H) {
    return img + ((size_t)job_tile0(j, H) + (size_t)nc * job_blocks(j) * (H / 64)) * WSLOT;
}

}  // namespace x3
}  // namespace bmp

using namespace bmp;
using namespace bmp::x3;

// debug: p = device buffer of 8 x n long long; every rowgemm3 launch after this call fills the next 8-slot record
extern "C" void bmp_debug_set_buffer_x3(void *p) { g_dbg = (long long *)p; g_dbg_slot = 0; }

// Bytes of workspace the tensor-core fp32 path needs (weight images + per-step temporaries [+ a one-step stash for
// inference]); 0 = shape not covered (the FFMA kernels of ggnn.cu run instead).
extern "C" size_t bmp_ggnn_x3_workspace_bytes(int mb, int n_atoms, int hidden, int n_edge, int n_steps, int inference) {
    if (!shape_ok(n_atoms, hidden, n_edge) || mb <= 0 || n_steps <= 0 || n_steps > BMP_MAX_STEPS) return 0;
    return Layout((long)mb * n_atoms, hidden, n_steps, inference != 0).total;
}

bool bmp_ggnn_x3_usable(int mb, int N, int H, int E, int T, const void *ws, size_t ws_bytes, const void *state_in, bool inference) {
    if (!ws || state_in || !shape_ok(N, H, E)) return false;
    return ws_bytes >= Layout((long)mb * N, H, T, inference).total;
}

int bmp_ggnn_forward_x3(const bmp_ggnn_fwd_t *a, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int H = a->hidden, N = a->n_atoms, T = a->n_steps, NC = H < 128 ? H : 128, hc = (H + 127) / 128, kb = H / 64;
    const long rows = (long)a->mb * N;
    const bool inference = !a->Hs;
    if (!inference && (!a->Ms || !a->Gs || !a->RSs)) { set_error("bmp_ggnn_forward: partial stash"); return BMP_EINVAL; }
    uint8_t *ws = (uint8_t *)(((uintptr_t)a->tc_workspace + 1023) & ~(uintptr_t)1023);
    const Layout L(rows, H, T, inference);
    int img_of[BMP_MAX_STEPS];
    image_plan(a, img_of);
    int rc;
    if (!a->tc_images_ready && (rc = pack_images(a, ws, L, img_of, st))) return rc;
    float *AH = reinterpret_cast<float *>(ws + L.tmp_off), *deg = reinterpret_cast<float *>(ws + L.deg_off);
    float *mini = reinterpret_cast<float *>(ws + L.mini_off);
    const size_t RH = (size_t)rows * H;
    auto Hs_at = [&](int t) { return inference ? mini + (size_t)(t & 1) * RH : a->Hs + (size_t)t * RH; };
    auto Ms_at = [&](int t) { return inference ? mini + 2 * RH : a->Ms + (size_t)t * RH; };
    auto Gs_at = [&](int t) { return inference ? mini + 3 * RH : a->Gs + (size_t)t * 3 * RH; };
    auto RS_at = [&](int t) { return inference ? mini + 6 * RH : a->RSs + (size_t)t * RH; };

    if (a->atoms) {
        if ((rc = bmp_embed_forward(a->atoms, a->embed_W, Hs_at(0), (int)rows, H, a->n_atom_types, stream))) return rc;
    } else {
        cudaMemcpyAsync(Hs_at(0), a->h_in, RH * sizeof(float), cudaMemcpyDeviceToDevice, st);
    }
    if (a->h0_out) cudaMemcpyAsync(a->h0_out, Hs_at(0), RH * sizeof(float), cudaMemcpyDeviceToDevice, st);

    const size_t agg_smem = (size_t)N * H * sizeof(float);
    cudaFuncSetAttribute(agg_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem);
    const int agg_grid = a->mb < 8 * sm_count() ? a->mb : 8 * sm_count();
    for (int t = 0; t < T; ++t) {
        const bool stf = a->stateful[t] != 0;
        const uint8_t *img = ws + (size_t)img_of[t] * L.img_bytes;
        const bmp_gru_t &G = a->gru[t];
        float *h = Hs_at(t), *hn = Hs_at(t + 1), *m = Ms_at(t), *g = Gs_at(t), *rs = RS_at(t);
        agg_fwd_kernel<<<agg_grid, 256, agg_smem, st>>>(a->adj, h, AH, t == 0 ? deg : nullptr, a->mb, N, H);
        count_launch();
        if ((rc = check_launch("agg_fwd_kernel"))) return rc;
        Args ga;
        // ---- message
        memset(&ga, 0, sizeof(ga));
        ga.rows = rows; ga.NC = NC; ga.njobs = hc;
        for (int nc = 0; nc < hc; ++nc) {
            Job &J = ga.job[nc];
            J.nblk = 4;
            for (int e = 0; e < 4; ++e) { J.A[e] = AH + (size_t)e * H; J.lda[e] = 4 * H; J.kt[e] = kb; }
            J.wimg = job_img(img, EPI_MSG, nc, H);
            J.epi = EPI_MSG;
            J.deg = deg; J.msg_b = a->msg_b[t] + (size_t)nc * 128 * 4;
            J.out0 = m + nc * 128; J.lo0 = H;
        }
        if ((rc = launch_gemm(ga, st))) return rc;
        // ---- reset (stateful) and update gates
        memset(&ga, 0, sizeof(ga));
        ga.rows = rows; ga.NC = NC;
        int nj = 0;
        for (int nc = 0; nc < hc; ++nc) {
            if (stf) {
                Job &J = ga.job[nj++];
                J.nblk = 2;
                J.A[0] = h; J.A[1] = m; J.lda[0] = J.lda[1] = H; J.kt[0] = J.kt[1] = kb;
                J.wimg = job_img(img, EPI_R, nc, H);
                J.epi = EPI_R;
                J.bias = G.b_Wr + nc * 128; J.bias2 = G.b_Ur + nc * 128;
                J.in0 = h + nc * 128; J.li0 = H;
                J.out0 = g + nc * 128; J.lo0 = 3 * H;
                J.out1 = rs + nc * 128; J.lo1 = H;
            }
            Job &J = ga.job[nj++];
            J.nblk = 2;
            J.A[0] = h; J.A[1] = m; J.lda[0] = J.lda[1] = H; J.kt[0] = J.kt[1] = kb;
            J.wimg = job_img(img, EPI_Z, nc, H);
            J.epi = EPI_Z;
            J.bias = G.b_Wz + nc * 128; J.bias2 = stf ? G.b_Uz + nc * 128 : nullptr;
            J.out0 = g + H + nc * 128; J.lo0 = 3 * H;
            if (!stf) { J.out1 = g + nc * 128; J.lo1 = 3 * H; J.out2 = rs + nc * 128; J.lo2 = H; }
        }
        ga.njobs = nj;
        if ((rc = launch_gemm(ga, st))) return rc;
        // ---- candidate and the new state
        memset(&ga, 0, sizeof(ga));
        ga.rows = rows; ga.NC = NC; ga.njobs = hc;
        for (int nc = 0; nc < hc; ++nc) {
            Job &J = ga.job[nc];
            J.nblk = stf ? 3 : 2;
            J.A[0] = h; J.A[1] = m; J.A[2] = rs; J.lda[0] = J.lda[1] = J.lda[2] = H; J.kt[0] = J.kt[1] = J.kt[2] = kb;
            J.wimg = job_img(img, EPI_HB, nc, H);
            J.epi = EPI_HB;
            J.bias = G.b_W + nc * 128; J.bias2 = stf ? G.b_U + nc * 128 : nullptr;
            J.in0 = g + H + nc * 128; J.li0 = 3 * H;
            J.in1 = stf ? h + nc * 128 : nullptr; J.li1 = H;
            J.out0 = g + 2 * H + nc * 128; J.lo0 = 3 * H;
            J.out1 = hn + nc * 128; J.lo1 = H;
            if (t == T - 1 && a->h_out) { J.out2 = a->h_out + nc * 128; J.lo2 = H; }
        }
        if ((rc = launch_gemm(ga, st))) return rc;
    }
    return BMP_OK;
}

// data part of the backward: Gs <- delta_r | delta_z | delta_h, Ps <- A_e^T dm, dHs[0] <- total gradient w.r.t. h_0
int bmp_ggnn_backward_x3(const bmp_ggnn_bwd_t *a, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int H = a->hidden, N = a->n_atoms, T = a->n_steps, NC = H < 128 ? H : 128, hc = (H + 127) / 128, kb = H / 64;
    const long rows = (long)a->mb * N;
    uint8_t *ws = (uint8_t *)(((uintptr_t)a->tc_workspace + 1023) & ~(uintptr_t)1023);
    const Layout L(rows, H, T, false);
    int img_of[BMP_MAX_STEPS];
    image_plan(a, img_of);
    int rc;
    if (!a->tc_images_ready && (rc = pack_images(a, ws, L, img_of, st))) return rc;
    const size_t RH = (size_t)rows * H;
    float *tmp = reinterpret_cast<float *>(ws + L.tmp_off);
    float *ds = tmp, *dhx = tmp + RH, *dm = tmp + 2 * RH;
    const size_t agg_smem = ((size_t)N * H + 4 * (size_t)N * (N + 1)) * sizeof(float);
    cudaFuncSetAttribute(agg_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem);
    const int agg_grid = a->mb < 8 * sm_count() ? a->mb : 8 * sm_count();
    for (int t = T - 1; t >= 0; --t) {
        const bool stf = a->stateful[t] != 0;
        const uint8_t *img = ws + (size_t)img_of[t] * L.img_bytes;
        float *Gt = a->Gs + (size_t)t * 3 * RH;
        const float *s = a->Hs + (size_t)t * RH;
        const float *g = a->dHs + (size_t)(t + 1) * RH;
        float *gout = a->dHs + (size_t)t * RH;
        const long nvec = rows * (H / 4);
        const int pgrid = (int)((nvec + 255) / 256 < 16L * sm_count() ? (nvec + 255) / 256 : 16L * sm_count());
        gate_bwd_kernel<<<pgrid, 256, 0, st>>>(g, Gt, stf ? s : nullptr, ds, rows, H);
        count_launch();
        if ((rc = check_launch("gate_bwd_kernel"))) return rc;
        Args ga;
        if (stf) {
            memset(&ga, 0, sizeof(ga));
            ga.rows = rows; ga.NC = NC; ga.njobs = hc;
            for (int nc = 0; nc < hc; ++nc) {
                Job &J = ga.job[nc];
                J.nblk = 1;
                J.A[0] = Gt + 2 * H; J.lda[0] = 3 * H; J.kt[0] = kb;
                J.wimg = job_img(img, EPI_Q, nc, H);
                J.epi = EPI_Q;
                J.in0 = s + nc * 128; J.li0 = H;
                J.in1 = Gt + nc * 128; J.li1 = 3 * H;
                J.in2 = ds + nc * 128; J.li2 = H;
                J.out0 = Gt + nc * 128; J.lo0 = 3 * H;
                J.out1 = ds + nc * 128; J.lo1 = H;
            }
            if ((rc = launch_gemm(ga, st))) return rc;
        }
        // ---- dh_x (+ ds) and dm
        memset(&ga, 0, sizeof(ga));
        ga.rows = rows; ga.NC = NC;
        int nj = 0;
        for (int nc = 0; nc < hc; ++nc)
            for (int which = 0; which < 2; ++which) {
                Job &J = ga.job[nj++];
                J.nblk = stf ? 3 : 2;
                J.A[0] = Gt + H; J.A[1] = Gt + 2 * H; J.A[2] = Gt;
                J.lda[0] = J.lda[1] = J.lda[2] = 3 * H; J.kt[0] = J.kt[1] = J.kt[2] = kb;
                J.wimg = job_img(img, which ? EPI_DM : EPI_DHX, nc, H);
                J.epi = which ? EPI_DM : EPI_DHX;
                if (!which && stf) { J.in0 = ds + nc * 128; J.li0 = H; }
                J.out0 = (which ? dm : dhx) + nc * 128; J.lo0 = H;
            }
        ga.njobs = nj;
        if ((rc = launch_gemm(ga, st))) return rc;
        agg_bwd_kernel<<<agg_grid, 256, agg_smem, st>>>(a->adj, dm, a->Ps + (size_t)t * 4 * RH, a->mb, N, H);
        count_launch();
        if ((rc = check_launch("agg_bwd_kernel"))) return rc;
        // ---- dHs[t] += dh_x + sum_e P_e W_e
        memset(&ga, 0, sizeof(ga));
        ga.rows = rows; ga.NC = NC; ga.njobs = hc;
        for (int nc = 0; nc < hc; ++nc) {
            Job &J = ga.job[nc];
            J.nblk = 4;
            for (int e = 0; e < 4; ++e) { J.A[e] = a->Ps + (size_t)t * 4 * RH + (size_t)e * H; J.lda[e] = 4 * H; J.kt[e] = kb; }
            J.wimg = job_img(img, EPI_DHMSG, nc, H);
            J.epi = EPI_DHMSG;
            J.in0 = dhx + nc * 128; J.li0 = H;
            J.in1 = gout + nc * 128; J.li1 = H;
            J.out0 = gout + nc * 128; J.lo0 = H;
        }
        if ((rc = launch_gemm(ga, st))) return rc;
    }
    return BMP_OK;
}
