// ggnn_tc256.cu -- fused GGNN encoder at hidden 256 on the 5th-gen tensor cores (BMP_MODE_BF16, forward only).
//
// BASELINE config D (GGNN hidden 256, T = 8, inference).  The H <= 128 kernel of ggnn_tc.cu keeps 4H TMEM columns and
// 160 KB of operand panels per tile, which does not exist at H = 256; this kernel re-slices the same step so that a tile
// (two padded molecules = 128 rows = one UMMA M) needs 3 x 64 KB of operand panels and 512 TMEM columns:
//   * message phase, 8 groups g = (bond-type pair p, 64-column block cb):
//       MMA-1  AH_g[(tp,i), c]  = sum_j A_{2p+tp}[i,j] h[j, 64cb + c]        per molecule, N = 64, B read MN-major
//       E1     AH_g -> two bf16 A panels (bond types 2p, 2p+1) in a 2-slot ring
//       MMA-2  m += AH_panel W_{e}[:, 64cb..]^T                               K slice of 64, two N = 128 halves
//     (E1 of group g+1 overlaps MMA-2 of group g; accumulators ping-pong in TMEM columns [256,512))
//   * E2  m + deg_e b_e -> bf16 panels, written over the adjacency (dead for the rest of the step; it is re-loaded for the
//     next step by one bulk copy from a per-CTA, L2-resident bf16 image dumped at tile start)
//   * gate phase: r -> TMEM [0,256), z -> TMEM [256,512); E3 forms r*h panels over the AH ring while z runs;
//     hbar = [h|m] W^T + (r*h) U^T overwrites r;  E4: h <- z*hbar + (1-z)*h (fp32 master state in registers).
// Weights stream from L2 as [256 n][16 k] slices (8 KB, un-swizzled core-matrix layout: one N = 256 UMMA each, so the A panel
// is read once per 256 output columns) through a 4-stage mbarrier ring -- 176 slices per stateful step.
// Warp roles as in ggnn_tc.cu: 16 epilogue warps (TMEM lane quarter = warp%4; 16-column chunk warp/4 of every 64-column
// block), one TMA producer warp, one MMA issuer warp.
// Replaces models/update/ggnn_update.py:31-63 / models/models/ggnn.py:72-106 at hidden 256 (forward).
#include "tc_common.cuh"

namespace bmp {
namespace tc256 {
using namespace bmp::tc;

constexpr int H = 256, KP = 4, EPW = 16, NE = 32 * EPW;
constexpr int TILE_BYTES = 256 * 16 * 2;              // weight slice: 256 rows (n) x 16 bf16 (k) = one UMMA k-step
constexpr int STAGES = 4;
constexpr int S_MSG = 16, S_GATE = 8, S_U = 4;        // 64-wide K slices per block (4 ring slices each)
constexpr int TILES_STATEFUL = 4 * (S_MSG + 3 * S_GATE + S_U), TILES_STATELESS = 4 * (S_MSG + 2 * S_GATE);
constexpr int OFF_H = 0, OFF_X = 4 * PANEL_BYTES, OFF_Y = 8 * PANEL_BYTES, OFF_W = 12 * PANEL_BYTES;
constexpr int OFF_BAR = OFF_W + STAGES * TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr uint32_t REGA = 0, REGB = 256;              // TMEM column regions
constexpr int ADJ_IMG_BYTES = 8 * ADJ_TILE_BYTES;     // staged adjacency of a tile: 64 KB
constexpr int MAX_CTAS = 160;

// un-swizzled K-major B slice [256 n][16 k]: core matrices (8 n x 16 B) contiguous; the two k halves 4096 B apart (LBO),
// consecutive 8-row groups 128 B apart (SBO)
__device__ __forceinline__ uint64_t desc_kmajor_plain(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

struct Args {
    int mb, N, T, n_types;
    const int32_t *atoms;
    const float *embed_W, *h_in;
    const void *adj;
    int adj_u8;
    const uint8_t *img[BMP_MAX_STEPS];
    const float *bias3[BMP_MAX_STEPS];
    const float *msg_b[BMP_MAX_STEPS];
    int stateful[BMP_MAX_STEPS];
    float *h_out, *h0_out;
    const int32_t *midx;         // optional (mb,): table-row indirection (see ggnn_tc.cu)
    uint8_t *scratch;            // gridDim.x x 64 KB: bf16 adjacency image of the CTA's current tile
    long long *dbg;              // optional: per step of CTA 0, the MMA lane's [total, weight wait, AH wait, x wait, rs wait, h wait] cycles
};

__global__ void __launch_bounds__(32 * (EPW + 2), 1) ggnn_tc256_kernel(const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = s32(smem);
    const uint32_t s_h = sbase + OFF_H, s_x = sbase + OFF_X, s_y = sbase + OFF_Y, s_w = sbase + OFF_W;
    const uint32_t s_bar = sbase + OFF_BAR;
    auto BAR = [&](int i) { return s_bar + 8u * i; };
    constexpr int B_FULL = 0, B_EMPTY = 4, B_HREADY = 8, B_ADJ = 9, B_D1 = 10, B_AHREADY = 14, B_AHFREE = 16, B_M = 18,
                  B_XREADY = 19, B_R = 20, B_RSREADY = 21, B_ZH = 22, NBAR = 23;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * NBAR + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (a.mb + 1) / 2;
    uint8_t *my_scratch = a.scratch + (size_t)blockIdx.x * ADJ_IMG_BYTES;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), 1); }
        mbar_init(BAR(B_HREADY), EPW);
        mbar_init(BAR(B_ADJ), 1);
        for (int i = 0; i < 4; ++i) mbar_init(BAR(B_D1 + i), 1);
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_AHREADY + i), EPW); mbar_init(BAR(B_AHFREE + i), 1); }
        mbar_init(BAR(B_M), 1);
        mbar_init(BAR(B_XREADY), EPW);
        mbar_init(BAR(B_R), 1);
        mbar_init(BAR(B_RSREADY), EPW);
        mbar_init(BAR(B_ZH), 1);
        fence_mbar_init();
    }
    if (warp == EPW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == EPW) {
        // ===================== TMA producer: weight tiles in consumption order + the per-step adjacency reload
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int t = 0; t < a.T; ++t, ++it) {
                    const int ntiles = a.stateful[t] ? TILES_STATEFUL : TILES_STATELESS;
                    const uint8_t *src = a.img[t];
                    for (int s = 0; s < ntiles; ++s) {
                        mbar_wait(BAR(B_EMPTY + stage), phase ^ 1);
                        mbar_expect_tx(BAR(B_FULL + stage), TILE_BYTES);
                        tma_bulk_g2s(s_w + stage * TILE_BYTES, src + (size_t)s * TILE_BYTES, TILE_BYTES, BAR(B_FULL + stage));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (t + 1 < a.T) {      // every MMA of step t is complete: the m panels are dead, bring the adjacency back
                        mbar_wait(BAR(B_ZH), it & 1);
                        asm volatile("fence.proxy.async;" ::: "memory");
                        mbar_expect_tx(BAR(B_ADJ), ADJ_IMG_BYTES);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            tma_bulk_g2s(s_x + c * PANEL_BYTES, my_scratch + c * PANEL_BYTES, PANEL_BYTES, BAR(B_ADJ));
                    }
                }
        }
    } else if (warp == EPW + 1) {
        // ===================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t ID_W = idesc(256, 0), ID_AH = idesc(64, 1);
            uint32_t stage = 0, phase = 0, it = 0, nadj = 0;
            long long wsum = 0, ahsum = 0;
            // one 64-wide K slice against all 256 output columns: four ring slices, one N = 256 UMMA each
            auto mma_kslice = [&](uint32_t a_addr, uint32_t dbase, bool first) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const long long w0 = a.dbg ? clock64() : 0;
                    mbar_wait(BAR(B_FULL + stage), phase);
                    if (a.dbg) wsum += clock64() - w0;
                    tc_fence_after();
                    tc_mma(tmem + dbase, desc_kmajor(a_addr + kk * 32), desc_kmajor_plain(s_w + stage * TILE_BYTES), ID_W, (first && kk == 0) ? 0u : 1u);
                    tc_commit(BAR(B_EMPTY + stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            };
            auto mma1 = [&](int g) {
                const int slot = g & 1, p = g >> 2, cb = g & 3;
                for (int mol = 0; mol < 2; ++mol) {
                    const uint32_t a_addr = s_x + (mol * 4 + 2 * p) * ADJ_TILE_BYTES;
                    const uint32_t b_addr = s_h + cb * PANEL_BYTES + mol * 64 * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma(tmem + REGB + slot * 128 + mol * 64, desc_kmajor(a_addr + k * 32),
                               desc_mnmajor(b_addr + k * 16 * 128), ID_AH, k ? 1u : 0u);
                    tc_commit(BAR(B_D1 + slot * 2 + mol));
                }
            };
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int t = 0; t < a.T; ++t, ++it) {
                    const uint32_t par = it & 1;
                    const bool stateful = a.stateful[t] != 0;
                    const long long c0 = a.dbg ? clock64() : 0;
                    mbar_wait(BAR(B_HREADY), par);
                    if (t > 0) { mbar_wait(BAR(B_ADJ), nadj & 1); ++nadj; }
                    tc_fence_after();
                    const long long c1 = a.dbg ? clock64() : 0;
                    // ---- message phase
                    mma1(0);
                    mma1(1);
                    for (int g = 0; g < 8; ++g) {
                        const int slot = g & 1;
                        const long long w0 = a.dbg ? clock64() : 0;
                        mbar_wait(BAR(B_AHREADY + slot), (g >> 1) & 1);
                        if (a.dbg) ahsum += clock64() - w0;
                        tc_fence_after();
                        for (int tp = 0; tp < 2; ++tp)
                            mma_kslice(s_y + (slot * 2 + tp) * PANEL_BYTES, REGA, g == 0 && tp == 0);
                        tc_commit(BAR(B_AHFREE + slot));
                        if (g + 2 < 8) mma1(g + 2);
                    }
                    tc_commit(BAR(B_M));
                    // ---- gate phase over x = [h | m]
                    // z, h part: needs only the state panels, runs while E2 forms the m panels (TMEM [256,512) is free after the last E1)
                    for (int kp = 0; kp < KP; ++kp) mma_kslice(s_h + kp * PANEL_BYTES, REGB, kp == 0);
                    const long long c2 = a.dbg ? clock64() : 0;
                    mbar_wait(BAR(B_XREADY), par);
                    tc_fence_after();
                    const long long c3 = a.dbg ? clock64() : 0;
                    auto xa = [&](int kp) { return kp < KP ? s_h + kp * PANEL_BYTES : s_x + (kp - KP) * PANEL_BYTES; };
                    if (stateful)
                        for (int kp = 0; kp < 2 * KP; ++kp) mma_kslice(xa(kp), REGA, kp == 0);        // r
                    tc_commit(BAR(B_R));
                    for (int kp = KP; kp < 2 * KP; ++kp) mma_kslice(xa(kp), REGB, false);             // z, m part (E3 overlaps)
                    const long long c4 = a.dbg ? clock64() : 0;
                    mbar_wait(BAR(B_RSREADY), par);         // r has been read, r*h panels are in place
                    tc_fence_after();
                    const long long c5 = a.dbg ? clock64() : 0;
                    for (int kp = 0; kp < 2 * KP; ++kp) mma_kslice(xa(kp), REGA, kp == 0);            // hbar (W part) overwrites r
                    if (stateful)
                        for (int kp = 0; kp < KP; ++kp) mma_kslice(s_y + kp * PANEL_BYTES, REGA, false);
                    tc_commit(BAR(B_ZH));
                    if (a.dbg && blockIdx.x == 0 && it < 64) {
                        long long *d = a.dbg + it * 8;
                        d[0] = clock64() - c0; d[1] = wsum; d[2] = ahsum; d[3] = c3 - c2; d[4] = c5 - c4; d[5] = c1 - c0; d[6] = c0;
                        wsum = ahsum = 0;
                    }
                }
        }
    } else {
        // ===================== epilogue warps 0..15
        const int q = warp & 3, cg = warp >> 2;
        const int row = 32 * q + lane;            // TMEM lane == tile row
        const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);
        const int molslot = row >> 6, atom = row & 63;
        float hreg[64];                           // fp32 master state: columns 64kb + 16cg + x at [16kb + x]
        float *stg = reinterpret_cast<float *>(smem + OFF_Y + warp * 2048);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int molg = tile * 2 + molslot;
            const bool live = molg < a.mb && atom < a.N;
            const long grow = (long)molg * a.N + atom;
            auto wrow = [&](int r) -> long {
                const int tr = 32 * q + r, mg = tile * 2 + (tr >> 6), at = tr & 63;
                return (mg < a.mb && at < a.N) ? (long)mg * a.N + at : -1L;
            };
            auto store_state = [&](float *base) {     // all 64 columns of this thread, 16 per pass (coalesced rows)
#pragma unroll
                for (int kb = 0; kb < KP; ++kb)
                    warp_store_rows<16>(stg, &hreg[16 * kb], lane, [&](int r) -> float * {
                        const long g = wrow(r);
                        return g >= 0 ? base + g * H + 64 * kb + 16 * cg : nullptr;
                    });
            };
            stage_adjacency<NE>(smem + OFF_X, a.adj, a.adj_u8, tile, a.mb, a.N, tid, a.midx);
            // ---- h_0: embedding gather (or h_in) -> fp32 registers
            {
                const float *src = nullptr;
                if (live) {
                    if (a.atoms) {
                        int id = __ldg(a.atoms + (a.midx ? (long)__ldg(a.midx + molg) * a.N + atom : grow));
                        id = id < 0 ? 0 : (id >= a.n_types ? a.n_types - 1 : id);
                        src = a.embed_W + (long)id * H;
                    } else {
                        src = a.h_in + grow * H;
                    }
                }
#pragma unroll
                for (int kb = 0; kb < KP; ++kb)
#pragma unroll
                    for (int c = 0; c < 16; c += 4) {
                        float4 v = src ? __ldg(reinterpret_cast<const float4 *>(src + 64 * kb + 16 * cg + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        hreg[16 * kb + c] = v.x; hreg[16 * kb + c + 1] = v.y; hreg[16 * kb + c + 2] = v.z; hreg[16 * kb + c + 3] = v.w;
                    }
            }
            if (a.h0_out) store_state(a.h0_out);
            auto store_h_operand = [&]() {
#pragma unroll
                for (int kb = 0; kb < KP; ++kb)
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const float *hv = &hreg[16 * kb + 8 * g];
                        uint4 pk = make_uint4(pack_bf16(hv[0], hv[1]), pack_bf16(hv[2], hv[3]), pack_bf16(hv[4], hv[5]), pack_bf16(hv[6], hv[7]));
                        *reinterpret_cast<uint4 *>(smem + OFF_H + kb * PANEL_BYTES + sw128(row, 16 * cg + 8 * g)) = pk;
                    }
            };
            store_h_operand();
            fence_proxy_async();
            asm volatile("bar.sync 1, %0;" ::"n"(NE));
            if (tid == 0 && a.T > 1) {      // bf16 image of the staged adjacency -> this CTA's scratch (re-loaded every later step)
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_bulk_s2g(my_scratch + c * PANEL_BYTES, s_x + c * PANEL_BYTES, PANEL_BYTES);
                bulk_commit();
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                __threadfence();
            }
            float deg[4];
            {
                float *dsh = reinterpret_cast<float *>(smem + OFF_Y);       // [4][128]
                {
                    const int e = cg;
                    float s0 = 0.f, s1 = 0.f;
                    const uint8_t *rowp = smem + OFF_X + (molslot * 4 + e) * ADJ_TILE_BYTES + atom * 128;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        uint4 u = *reinterpret_cast<const uint4 *>(rowp + ((ch ^ (atom & 7)) << 4));
                        const __nv_bfloat162 *b2 = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
                        for (int x = 0; x < 4; ++x) { float2 f = __bfloat1622float2(b2[x]); s0 += f.x; s1 += f.y; }
                    }
                    dsh[e * 128 + row] = s0 + s1;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NE));
#pragma unroll
                for (int e = 0; e < 4; ++e) deg[e] = dsh[e * 128 + row];
                asm volatile("bar.sync 1, %0;" ::"n"(NE));
            }
            warp_arrive(BAR(B_HREADY), lane);

            for (int t = 0; t < a.T; ++t, ++it) {
                const uint32_t par = it & 1;
                const bool stateful = a.stateful[t] != 0;
                const float *b3 = a.bias3[t];
                if (t == a.T - 1 && tile + (int)gridDim.x < n_tiles) {
                    prefetch_adjacency_l2<NE>(a.adj, a.adj_u8, tile + gridDim.x, a.mb, a.N, tid, a.midx);
                    if (a.atoms && !a.midx && tid < 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.atoms + (long)(tile + gridDim.x) * 2 * a.N + tid * 32));
                }
                uint32_t w[16];
                // ---- E1: AH accumulators of group g -> the two bf16 A panels of ring slot g&1
                for (int g = 0; g < 8; ++g) {
                    const int slot = g & 1;
                    mbar_wait(BAR(B_AHFREE + slot), ((g >> 1) + 1) & 1);     // MMA-2 of group g-2 has consumed the slot
                    for (int mol = 0; mol < 2; ++mol) {
                        mbar_wait(BAR(B_D1 + slot * 2 + mol), (g >> 1) & 1);
                        tc_fence_after();
                        tc_ld16(t_lane + REGB + slot * 128 + mol * 64 + 16 * cg, w);
                        tc_wait_ld();
                        const int orow = mol * 64 + atom;       // TMEM lane half (molslot) = bond type within the pair
#pragma unroll
                        for (int gg = 0; gg < 2; ++gg) {
                            uint4 pk = make_uint4(pack_bf16(__uint_as_float(w[8 * gg]), __uint_as_float(w[8 * gg + 1])),
                                                  pack_bf16(__uint_as_float(w[8 * gg + 2]), __uint_as_float(w[8 * gg + 3])),
                                                  pack_bf16(__uint_as_float(w[8 * gg + 4]), __uint_as_float(w[8 * gg + 5])),
                                                  pack_bf16(__uint_as_float(w[8 * gg + 6]), __uint_as_float(w[8 * gg + 7])));
                            *reinterpret_cast<uint4 *>(smem + OFF_Y + (slot * 2 + molslot) * PANEL_BYTES + sw128(orow, 16 * cg + 8 * gg)) = pk;
                        }
                    }
                    warp_arrive(BAR(B_AHREADY + slot), lane);
                }
                // ---- E2: message m (+ bias through the degrees) -> bf16 panels over the adjacency
                mbar_wait(BAR(B_M), par);
                tc_fence_after();
                {
                    const float4 *mb4 = reinterpret_cast<const float4 *>(a.msg_b[t]);
#pragma unroll
                    for (int kb = 0; kb < KP; ++kb) {
                        tc_ld16(t_lane + REGA + 64 * kb + 16 * cg, w);
                        tc_wait_ld();
                        float m[16];
#pragma unroll
                        for (int x = 0; x < 16; ++x) {
                            const float4 b4 = __ldg(mb4 + 64 * kb + 16 * cg + x);
                            m[x] = __uint_as_float(w[x]) + deg[0] * b4.x + deg[1] * b4.y + deg[2] * b4.z + deg[3] * b4.w;
                        }
#pragma unroll
                        for (int gg = 0; gg < 2; ++gg) {
                            uint4 pk = make_uint4(pack_bf16(m[8 * gg], m[8 * gg + 1]), pack_bf16(m[8 * gg + 2], m[8 * gg + 3]),
                                                  pack_bf16(m[8 * gg + 4], m[8 * gg + 5]), pack_bf16(m[8 * gg + 6], m[8 * gg + 7]));
                            *reinterpret_cast<uint4 *>(smem + OFF_X + kb * PANEL_BYTES + sw128(row, 16 * cg + 8 * gg)) = pk;
                        }
                    }
                }
                warp_arrive(BAR(B_XREADY), lane);
                // ---- E3: reset gate, r*h -> bf16 panels over the AH ring
                mbar_wait(BAR(B_R), par);
                tc_fence_after();
                if (stateful) {
#pragma unroll
                    for (int kb = 0; kb < KP; ++kb) {
                        tc_ld16(t_lane + REGA + 64 * kb + 16 * cg, w);
                        tc_wait_ld();
                        float rs[16];
#pragma unroll
                        for (int x = 0; x < 16; x += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(b3 + 64 * kb + 16 * cg + x));
                            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                            for (int y = 0; y < 4; ++y)
                                rs[x + y] = sigmoid_fast(__uint_as_float(w[x + y]) + bb[y]) * hreg[16 * kb + x + y];
                        }
#pragma unroll
                        for (int gg = 0; gg < 2; ++gg) {
                            uint4 pk = make_uint4(pack_bf16(rs[8 * gg], rs[8 * gg + 1]), pack_bf16(rs[8 * gg + 2], rs[8 * gg + 3]),
                                                  pack_bf16(rs[8 * gg + 4], rs[8 * gg + 5]), pack_bf16(rs[8 * gg + 6], rs[8 * gg + 7]));
                            *reinterpret_cast<uint4 *>(smem + OFF_Y + kb * PANEL_BYTES + sw128(row, 16 * cg + 8 * gg)) = pk;
                        }
                    }
                }
                warp_arrive(BAR(B_RSREADY), lane);
                // ---- E4: update gate + candidate -> new state
                mbar_wait(BAR(B_ZH), par);
                tc_fence_after();
#pragma unroll
                for (int kb = 0; kb < KP; ++kb) {
                    uint32_t wh[16];
                    tc_ld16(t_lane + REGB + 64 * kb + 16 * cg, w);
                    tc_ld16(t_lane + REGA + 64 * kb + 16 * cg, wh);
                    tc_wait_ld();
#pragma unroll
                    for (int x = 0; x < 16; x += 4) {
                        const float4 bz4 = __ldg(reinterpret_cast<const float4 *>(b3 + H + 64 * kb + 16 * cg + x));
                        const float4 bh4 = __ldg(reinterpret_cast<const float4 *>(b3 + 2 * H + 64 * kb + 16 * cg + x));
                        const float bz[4] = {bz4.x, bz4.y, bz4.z, bz4.w}, bh[4] = {bh4.x, bh4.y, bh4.z, bh4.w};
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            const float z = sigmoid_fast(__uint_as_float(w[x + y]) + bz[y]);
                            const float hb = tanh_fast(__uint_as_float(wh[x + y]) + bh[y]);
                            float &hv = hreg[16 * kb + x + y];
                            hv = stateful ? fmaf(z, hb - hv, hv) : z * hb;
                        }
                    }
                }
                if (t + 1 < a.T) {
                    store_h_operand();
                    warp_arrive(BAR(B_HREADY), lane);
                } else {
                    if (a.h_out) store_state(a.h_out);      // staging = the AH ring, idle after the last MMA
                    tc_fence_before();
                    asm volatile("bar.sync 1, %0;" ::"n"(NE));   // everyone done with this tile's smem/TMEM
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPW + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ---------------------------------------------------------------- weight packing
// bf16 B-operand slices [256 n][16 k] of one step in consumption order (four per 64-wide K slice):
//   [MMA-2: group g = (p, cb) x bond type tp] [z: h blocks] [r: K block] [z: m blocks] [hbar] [U]   (stateless: no r, no U)
struct PackArgs {
    int stateful;
    const float *msg_W;
    bmp_gru_t g;
    uint8_t *img;
    float *bias3;
};

__global__ void pack256_kernel(const PackArgs p) {
    const int nslices = p.stateful ? TILES_STATEFUL : TILES_STATELESS;
    const long total = (long)nslices * 256 * 16;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int k = idx & 15, n = (idx >> 4) & 255, slice = (int)(idx >> 12);
        const int ks = slice >> 2, kin = (slice & 3) * 16 + k;       // 64-wide K slice, column inside it
        float w;
        if (ks < S_MSG) {
            const int g = ks >> 1, tp = ks & 1, e = 2 * (g >> 2) + tp, cp = (g & 3) * 64 + kin;
            w = p.msg_W[((long)n * 4 + e) * H + cp];
        } else {
            // gate K slices in issue order: z (h part: 4) | r (8, stateful only) | z (m part: 4) | hbar (8) | U (4, stateful only)
            int j = ks - S_MSG, g, kp;
            if (j < KP) { g = 1; kp = j; }
            else {
                j -= KP;
                if (p.stateful && j < 2 * KP) { g = 0; kp = j; }
                else {
                    if (p.stateful) j -= 2 * KP;
                    if (j < KP) { g = 1; kp = KP + j; }
                    else if (j < 3 * KP) { g = 2; kp = j - KP; }
                    else { g = 3; kp = j - 3 * KP; }
                }
            }
            if (g < 3) {
                const int K = kp * 64 + kin;
                const float *W = g == 0 ? p.g.W_r : (g == 1 ? p.g.W_z : p.g.W);
                w = W[(long)n * 2 * H + K];
                if (p.stateful && K < H && g < 2) w += (g == 0 ? p.g.U_r : p.g.U_z)[(long)n * H + K];
            } else {
                w = p.g.U[(long)n * H + kp * 64 + kin];
            }
        }
        const uint32_t off = (uint32_t)(k >> 3) * 4096u + (uint32_t)n * 16u + ((uint32_t)k & 7u) * 2u;
        *reinterpret_cast<__nv_bfloat16 *>(p.img + (size_t)slice * TILE_BYTES + off) = __float2bfloat16_rn(w);
    }
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < H; c += blockDim.x) {
            p.bias3[c] = p.stateful ? p.g.b_Wr[c] + p.g.b_Ur[c] : 0.f;
            p.bias3[H + c] = p.g.b_Wz[c] + (p.stateful ? p.g.b_Uz[c] : 0.f);
            p.bias3[2 * H + c] = p.g.b_W[c] + (p.stateful ? p.g.b_U[c] : 0.f);
        }
}

static size_t image_bytes() { return (size_t)TILES_STATEFUL * TILE_BYTES + 3 * H * sizeof(float) + 256; }

}  // namespace tc256
}  // namespace bmp

using namespace bmp;

extern long long *g_tc_dbg_fwd;                        // ggnn_tc.cu (bmp_debug_set_buffer_fwd)

size_t bmp_ggnn_tc256_workspace_bytes(int n_steps) {
    return tc256::image_bytes() * (size_t)n_steps + (size_t)tc256::MAX_CTAS * tc256::ADJ_IMG_BYTES + 4096;
}

int bmp_ggnn_forward_tc256(const bmp_ggnn_fwd_t *a, void *stream) {
    const int T = a->n_steps;
    if (a->n_edge != 4) { set_error("BMP_MODE_BF16: n_edge=%d not supported (4)", a->n_edge); return BMP_ESHAPE; }
    if (a->mb <= 0 || T <= 0 || T > BMP_MAX_STEPS || a->n_atoms <= 0 || a->n_atoms > BMP_MAX_ATOMS) {
        set_error("BMP_MODE_BF16: bad shape mb=%d T=%d N=%d", a->mb, T, a->n_atoms);
        return BMP_ESHAPE;
    }
    if (a->state_in) { set_error("BMP_MODE_BF16: an external GRU state is not supported (state must equal h)"); return BMP_ESHAPE; }
    if (a->Hs || a->Ms || a->Gs || a->RSs || a->stash2) {
        set_error("BMP_MODE_BF16: hidden 256 is forward-only (no training stash); train in BMP_MODE_F32");
        return BMP_ESHAPE;
    }
    if (!a->tc_workspace || a->tc_workspace_bytes < bmp_ggnn_tc256_workspace_bytes(T)) {
        set_error("BMP_MODE_BF16: tc_workspace of >= %zu bytes required", bmp_ggnn_tc256_workspace_bytes(T));
        return BMP_EINVAL;
    }
    if (!aligned16({a->h_in, a->embed_W, a->adj, a->h_out, a->h0_out, a->tc_workspace})) {
        set_error("BMP_MODE_BF16: buffers must be 16-byte aligned");
        return BMP_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    tc256::Args k = {};
    k.mb = a->mb; k.N = a->n_atoms; k.T = T; k.n_types = a->n_atom_types;
    k.atoms = a->atoms; k.embed_W = a->embed_W; k.h_in = a->h_in; k.adj = a->adj; k.adj_u8 = a->adj_u8;
    k.h_out = a->h_out; k.h0_out = a->h0_out;
    k.midx = a->mol_index;
    if (a->mol_index && !a->atoms) { set_error("BMP_MODE_BF16: mol_index needs atom ids (a drug table), not h_in"); return BMP_EINVAL; }
    k.dbg = g_tc_dbg_fwd;
    uint8_t *ws = (uint8_t *)(((uintptr_t)a->tc_workspace + 1023) & ~(uintptr_t)1023);
    const size_t ib = tc256::image_bytes();
    k.scratch = ws;
    uint8_t *imgs = ws + (size_t)tc256::MAX_CTAS * tc256::ADJ_IMG_BYTES;
    int n_img = 0;
    for (int t = 0; t < T; ++t) {
        if (!a->msg_W[t] || !a->msg_b[t] || !a->gru[t].W_z || !a->gru[t].W) { set_error("BMP_MODE_BF16: null parameter at step %d", t); return BMP_EINVAL; }
        if (((uintptr_t)a->msg_b[t]) & 15) { set_error("BMP_MODE_BF16: msg_b must be 16-byte aligned"); return BMP_EINVAL; }
        int found = -1;
        for (int u = 0; u < t; ++u)
            if (a->msg_W[u] == a->msg_W[t] && same_gru(a->gru[u], a->gru[t]) &&
                (a->stateful[u] != 0) == (a->stateful[t] != 0)) { found = u; break; }
        k.stateful[t] = a->stateful[t] != 0;
        k.msg_b[t] = a->msg_b[t];
        if (found >= 0) { k.img[t] = k.img[found]; k.bias3[t] = k.bias3[found]; continue; }
        uint8_t *img = imgs + (size_t)n_img * ib;
        float *bias3 = reinterpret_cast<float *>(img + (size_t)tc256::TILES_STATEFUL * tc256::TILE_BYTES);
        tc256::PackArgs p;
        p.stateful = k.stateful[t]; p.msg_W = a->msg_W[t]; p.g = a->gru[t]; p.img = img; p.bias3 = bias3;
        if (!a->tc_images_ready) {
            tc256::pack256_kernel<<<128, 256, 0, st>>>(p);
            count_launch();
        }
        k.img[t] = img; k.bias3[t] = bias3;
        ++n_img;
    }
    int rc = check_launch("pack256_kernel");
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > tc256::MAX_CTAS) sms = tc256::MAX_CTAS;
    const int n_tiles = (a->mb + 1) / 2;
    int grid = n_tiles < sms ? n_tiles : sms;
    if (const char *e = getenv("BMP_TC256_GRID")) { const int g = atoi(e); if (g > 0 && g < grid) grid = g; }   // experiments only
    cudaFuncSetAttribute(tc256::ggnn_tc256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc256::SMEM_BYTES);
    ProfScope prof(BMP_PROF_GGNN_FWD, st);
    tc256::ggnn_tc256_kernel<<<grid, 32 * (tc256::EPW + 2), tc256::SMEM_BYTES, st>>>(k);
    count_launch();
    return check_launch("ggnn_tc256_kernel");
}
