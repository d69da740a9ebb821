// bimpm.cu -- BiMPM matching co-attention (models/coattention/bimpm.py:17-197; `--attn bimpm`, train_binary.py:253-256),
// forward and backward, fp32, one CTA per drug pair.  An API-completeness kernel, not a tuned one.
//
// Per pair, X1 (N1 x H), X2 (N2 x H), three weights W (K x H), K = head, e = 1e-5 (F.normalize), E = 1e-4 (div_with_small_value):
//   1. max-pooling matching   S_k[i,j] = cos(W_k o X1_i, W_k o X2_j);  m1[i,k] = max_j S_k,  m2[j,k] = max_i S_k          (:84-108,:136-146)
//   2. attention              att[i,j] = cos(X1_i, X2_j)                                                                  (:110-123,:148-151)
//      attentive matching     M2_i = sum_j att[i,j] X2_j / max(sum_j att[i,j], E) (M1_j likewise);  r = match(X1, M2, W_mean) (:162-172)
//      max-attentive matching Z2[i,c] = max_j att[i,j] X2[j,c] (Z1 likewise);                      r = match(X1, Z2, W_attmax) (:174-186)
//      with match(X, V, W)[i,k] = cos(W_k o X_i, W_0 o V_i)  -- the reference multiplies (head x hidden) by (hidden x head) and keeps
//      column 0 (:79-81), so perspective k of the first operand meets perspective 0 of the second; reproduced as is.
//   3. out_1 = sum_i [m1 | r_mean1 | r_attmax1][i, :],  out_2 likewise over j  (aggr = F.sum, :188-197); 3K columns each.
// F.max sends the gradient to every position equal to the maximum; F.maximum(d, E) passes it where d >= E.
// Intermediates live in a per-CTA scratch area in global memory (L2-resident), parameter gradients accumulate in shared memory
// and are flushed once per CTA with atomics.
#include "common.cuh"

namespace bmp {
namespace bimpm {

constexpr float EPS_N = 1e-5f, EPS_D = 1e-4f;
constexpr int NT = 256;

struct Args {
    int mb, N1, N2, H, K;
    const float *X1, *X2;            // (mb,N1,H), (mb,N2,H)
    const float *W[3];               // max_pooling_W, att_mean_W, att_max_W : (K,H)
    float *out1, *out2;              // (mb,3K)
    const float *G1, *G2;            // backward: upstream gradients (mb,3K)
    float *dX1, *dX2;                // backward: (mb,N1,H), (mb,N2,H)  (overwritten)
    float *dW[3];                    // backward: accumulated (+=)
    float *scratch;                  // gridDim.x * scratch_floats
    long scratch_floats;
};

__host__ __device__ inline long scratch_floats(int N1, int N2, int H, int K) {
    const long a = (long)N1 * N2, r1 = (long)N1 * H, r2 = (long)N2 * H, k1 = (long)N1 * K, k2 = (long)N2 * K;
    // att, S | xn1 xn2 den2 den1 | M2 M1 Z2 Z1 | nm1 nm2 m1 m2 | np(4) r(4) nq(4) | dAtt dS | dM2 dM1 dZ2 dZ1 | T1 T2 dden2 dden1
    return 2 * a + 2L * (N1 + N2) + 2 * (r1 + r2) + 2 * (k1 + k2) + 4 * (k1 + k2) + 2L * (N1 + N2) + 2 * a + 2 * (r1 + r2) + 2L * (N1 + N2) + 64;
}

struct Scratch {
    float *att, *S, *xn1, *xn2, *den2, *den1, *M2, *M1, *Z2, *Z1, *nm1, *nm2, *m1, *m2;
    float *np[4], *r[4], *nq[4];     // matchings: 0 = (X1, M2, W_mean), 1 = (X2, M1, W_mean), 2 = (X1, Z2, W_attmax), 3 = (X2, Z1, W_attmax)
    float *dAtt, *dS, *dM2, *dM1, *dZ2, *dZ1, *T1, *T2, *dden2, *dden1;
    __device__ void carve(float *p, int N1, int N2, int H, int K) {
        const long a = (long)N1 * N2, r1 = (long)N1 * H, r2 = (long)N2 * H, k1 = (long)N1 * K, k2 = (long)N2 * K;
        att = p; p += a; S = p; p += a;
        xn1 = p; p += N1; xn2 = p; p += N2; den2 = p; p += N1; den1 = p; p += N2;
        M2 = p; p += r1; M1 = p; p += r2; Z2 = p; p += r1; Z1 = p; p += r2;
        nm1 = p; p += k1; nm2 = p; p += k2; m1 = p; p += k1; m2 = p; p += k2;
        for (int t = 0; t < 4; ++t) { const long kk = (t & 1) ? k2 : k1; np[t] = p; p += kk; r[t] = p; p += kk; }
        for (int t = 0; t < 4; ++t) { nq[t] = p; p += (t & 1) ? N2 : N1; }
        dAtt = p; p += a; dS = p; p += a;
        dM2 = p; p += r1; dM1 = p; p += r2; dZ2 = p; p += r1; dZ1 = p; p += r2;
        T1 = p; p += N1; T2 = p; p += N2; dden2 = p; p += N1; dden1 = p; p += N2;
    }
};

#define FOR(idx, n) for (long idx = threadIdx.x; idx < (long)(n); idx += NT)

// everything the forward computes, into the scratch area (the backward re-runs it)
__device__ void forward_pass(const Args &a, const Scratch &s, const float *X1, const float *X2) {
    const int N1 = a.N1, N2 = a.N2, H = a.H, K = a.K;
    const float *Wm = a.W[0], *Wa = a.W[1], *Wx = a.W[2];
    // F1: row norms of X and of every perspective W_k o X
    FOR(i, N1) { float q = 0.f; for (int c = 0; c < H; ++c) q = fmaf(X1[i * H + c], X1[i * H + c], q); s.xn1[i] = sqrtf(q); }
    FOR(j, N2) { float q = 0.f; for (int c = 0; c < H; ++c) q = fmaf(X2[j * H + c], X2[j * H + c], q); s.xn2[j] = sqrtf(q); }
    auto persp_norms = [&](const float *X, int N, const float *W, float *dst) {
        FOR(idx, (long)N * K) {
            const int i = idx / K, k = idx - i * K;
            float q = 0.f;
            for (int c = 0; c < H; ++c) { const float p = W[k * H + c] * X[i * H + c]; q = fmaf(p, p, q); }
            dst[idx] = sqrtf(q);
        }
    };
    persp_norms(X1, N1, Wm, s.nm1); persp_norms(X2, N2, Wm, s.nm2);
    persp_norms(X1, N1, Wa, s.np[0]); persp_norms(X2, N2, Wa, s.np[1]);
    persp_norms(X1, N1, Wx, s.np[2]); persp_norms(X2, N2, Wx, s.np[3]);
    __syncthreads();
    // F2: cosine attention
    FOR(idx, (long)N1 * N2) {
        const int i = idx / N2, j = idx - i * N2;
        float d = 0.f;
        for (int c = 0; c < H; ++c) d = fmaf(X1[i * H + c], X2[j * H + c], d);
        s.att[idx] = d / ((s.xn1[i] + EPS_N) * (s.xn2[j] + EPS_N));
    }
    __syncthreads();
    // F3: attention sums, attentive means, max-attentive vectors
    FOR(i, N1) { float q = 0.f; for (int j = 0; j < N2; ++j) q += s.att[i * N2 + j]; s.den2[i] = q; }
    FOR(j, N2) { float q = 0.f; for (int i = 0; i < N1; ++i) q += s.att[i * N2 + j]; s.den1[j] = q; }
    __syncthreads();
    FOR(idx, (long)N1 * H) {
        const int i = idx / H, c = idx - i * H;
        float sum = 0.f, mx = -INFINITY;
        for (int j = 0; j < N2; ++j) { const float v = s.att[i * N2 + j] * X2[j * H + c]; sum += v; mx = fmaxf(mx, v); }
        s.M2[idx] = sum / fmaxf(s.den2[i], EPS_D);
        s.Z2[idx] = mx;
    }
    FOR(idx, (long)N2 * H) {
        const int j = idx / H, c = idx - j * H;
        float sum = 0.f, mx = -INFINITY;
        for (int i = 0; i < N1; ++i) { const float v = s.att[i * N2 + j] * X1[i * H + c]; sum += v; mx = fmaxf(mx, v); }
        s.M1[idx] = sum / fmaxf(s.den1[j], EPS_D);
        s.Z1[idx] = mx;
    }
    __syncthreads();
    // F5: the four column-0 matchings
    for (int t = 0; t < 4; ++t) {
        const float *X = (t & 1) ? X2 : X1, *V = t == 0 ? s.M2 : (t == 1 ? s.M1 : (t == 2 ? s.Z2 : s.Z1)), *W = t < 2 ? Wa : Wx;
        const int N = (t & 1) ? N2 : N1;
        FOR(i, N) { float q = 0.f; for (int c = 0; c < H; ++c) { const float p = W[c] * V[i * H + c]; q = fmaf(p, p, q); } s.nq[t][i] = sqrtf(q); }
        __syncthreads();
        FOR(idx, (long)N * K) {
            const int i = idx / K, k = idx - i * K;
            float d = 0.f;
            for (int c = 0; c < H; ++c) d = fmaf(W[k * H + c] * X[i * H + c], W[c] * V[i * H + c], d);
            s.r[t][idx] = d / ((s.np[t][idx] + EPS_N) * (s.nq[t][i] + EPS_N));
        }
    }
    __syncthreads();
}

// S_k for the max-pooling matching into s.S
__device__ void maxpool_sim(const Args &a, const Scratch &s, const float *X1, const float *X2, int k) {
    const int N1 = a.N1, N2 = a.N2, H = a.H, K = a.K;
    const float *w = a.W[0] + (long)k * H;
    FOR(idx, (long)N1 * N2) {
        const int i = idx / N2, j = idx - i * N2;
        float d = 0.f;
        for (int c = 0; c < H; ++c) d = fmaf(w[c] * X1[i * H + c], w[c] * X2[j * H + c], d);
        s.S[idx] = d / ((s.nm1[i * K + k] + EPS_N) * (s.nm2[j * K + k] + EPS_N));
    }
    __syncthreads();
}

template <bool BWD>
__global__ void __launch_bounds__(NT) bimpm_kernel(const Args a) {
    extern __shared__ float sm_dw[];          // BWD: 3 x K x H parameter-gradient accumulators
    const int N1 = a.N1, N2 = a.N2, H = a.H, K = a.K;
    Scratch s;
    s.carve(a.scratch + (long)blockIdx.x * a.scratch_floats, N1, N2, H, K);
    if (BWD) {
        FOR(idx, 3L * K * H) sm_dw[idx] = 0.f;
        __syncthreads();
    }
    for (int b = blockIdx.x; b < a.mb; b += gridDim.x) {
        const float *X1 = a.X1 + (long)b * N1 * H, *X2 = a.X2 + (long)b * N2 * H;
        forward_pass(a, s, X1, X2);
        // F4: max-pooling matching, one perspective at a time
        for (int k = 0; k < K; ++k) {
            maxpool_sim(a, s, X1, X2, k);
            FOR(i, N1) { float mx = -INFINITY; for (int j = 0; j < N2; ++j) mx = fmaxf(mx, s.S[i * N2 + j]); s.m1[i * K + k] = mx; }
            FOR(j, N2) { float mx = -INFINITY; for (int i = 0; i < N1; ++i) mx = fmaxf(mx, s.S[i * N2 + j]); s.m2[j * K + k] = mx; }
            __syncthreads();
        }
        if (!BWD) {
            FOR(idx, 3L * K) {
                const int blk = idx / K, k = idx - blk * K;
                const float *src1 = blk == 0 ? s.m1 : (blk == 1 ? s.r[0] : s.r[2]);
                const float *src2 = blk == 0 ? s.m2 : (blk == 1 ? s.r[1] : s.r[3]);
                float q1 = 0.f, q2 = 0.f;
                for (int i = 0; i < N1; ++i) q1 += src1[i * K + k];
                for (int j = 0; j < N2; ++j) q2 += src2[j * K + k];
                a.out1[(long)b * 3 * K + idx] = q1;
                a.out2[(long)b * 3 * K + idx] = q2;
            }
            __syncthreads();
            continue;
        }
        // ================================================================ backward
        const float *G1 = a.G1 + (long)b * 3 * K, *G2 = a.G2 + (long)b * 3 * K;
        float *dX1 = a.dX1 + (long)b * N1 * H, *dX2 = a.dX2 + (long)b * N2 * H;
        FOR(idx, (long)N1 * H) { dX1[idx] = 0.f; s.dM2[idx] = 0.f; s.dZ2[idx] = 0.f; }
        FOR(idx, (long)N2 * H) { dX2[idx] = 0.f; s.dM1[idx] = 0.f; s.dZ1[idx] = 0.f; }
        __syncthreads();
        // B5: the four matchings.  r_k = a_k . b with a_k = normalize(W_k o x), b = normalize(W_0 o v):
        //   d(W_k o x) = g_k s_k (b - p_k r_k / n_k),   d(W_0 o v) = s_q (sum_k g_k a_k - q R / n_q),  R = sum_k g_k r_k
        for (int t = 0; t < 4; ++t) {
            const float *X = (t & 1) ? X2 : X1, *V = t == 0 ? s.M2 : (t == 1 ? s.M1 : (t == 2 ? s.Z2 : s.Z1));
            const float *W = a.W[t < 2 ? 1 : 2], *G = ((t & 1) ? G2 : G1) + (t < 2 ? K : 2 * K);
            float *dX = (t & 1) ? dX2 : dX1, *dV = t == 0 ? s.dM2 : (t == 1 ? s.dM1 : (t == 2 ? s.dZ2 : s.dZ1));
            float *dWs = sm_dw + (long)(t < 2 ? 1 : 2) * K * H;
            const int N = (t & 1) ? N2 : N1;
            FOR(idx, (long)N * H) {
                const int i = idx / H, c = idx - i * H;
                const float x = X[idx], v = V[idx], nq = s.nq[t][i], sq = 1.f / (nq + EPS_N);
                const float q = W[c] * v, bb = q * sq;
                float D = 0.f, R = 0.f, dx = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float g = G[k], nk = s.np[t][i * K + k], sk = 1.f / (nk + EPS_N), r = s.r[t][i * K + k];
                    const float p = W[k * H + c] * x;
                    const float dp = nk > 0.f ? g * sk * (bb - p * r / nk) : 0.f;
                    dx = fmaf(W[k * H + c], dp, dx);
                    atomicAdd(dWs + k * H + c, dp * x);
                    D = fmaf(g, p * sk, D);
                    R = fmaf(g, r, R);
                }
                const float dq = nq > 0.f ? sq * (D - q * R / nq) : 0.f;
                dX[idx] += dx;
                dV[idx] += W[c] * dq;
                atomicAdd(dWs + c, dq * v);
            }
            __syncthreads();
        }
        // B3: through the attentive means / max-attentive vectors to att and X
        FOR(i, N1) {      // d den2[i] (only where den2 >= E) ; dM2 <- d num2 = dM2 / max(den2, E)
            const float dd = fmaxf(s.den2[i], EPS_D);
            float q = 0.f;
            for (int c = 0; c < H; ++c) { q = fmaf(s.dM2[i * H + c], s.M2[i * H + c], q); s.dM2[i * H + c] /= dd; }
            s.dden2[i] = s.den2[i] >= EPS_D ? -q / dd : 0.f;
        }
        FOR(j, N2) {
            const float dd = fmaxf(s.den1[j], EPS_D);
            float q = 0.f;
            for (int c = 0; c < H; ++c) { q = fmaf(s.dM1[j * H + c], s.M1[j * H + c], q); s.dM1[j * H + c] /= dd; }
            s.dden1[j] = s.den1[j] >= EPS_D ? -q / dd : 0.f;
        }
        __syncthreads();
        FOR(idx, (long)N1 * N2) {
            const int i = idx / N2, j = idx - i * N2;
            const float at = s.att[idx];
            float d = s.dden2[i] + s.dden1[j];
            for (int c = 0; c < H; ++c) {
                const float x1 = X1[i * H + c], x2 = X2[j * H + c];
                d = fmaf(s.dM2[i * H + c], x2, d);
                d = fmaf(s.dM1[j * H + c], x1, d);
                if (at * x2 == s.Z2[i * H + c]) d = fmaf(s.dZ2[i * H + c], x2, d);
                if (at * x1 == s.Z1[j * H + c]) d = fmaf(s.dZ1[j * H + c], x1, d);
            }
            s.dAtt[idx] = d;
        }
        FOR(idx, (long)N2 * H) {
            const int j = idx / H, c = idx - j * H;
            const float x2 = X2[idx];
            float d = 0.f;
            for (int i = 0; i < N1; ++i) {
                const float at = s.att[i * N2 + j];
                d = fmaf(at, s.dM2[i * H + c], d);
                if (at * x2 == s.Z2[i * H + c]) d = fmaf(at, s.dZ2[i * H + c], d);
            }
            dX2[idx] += d;
        }
        FOR(idx, (long)N1 * H) {
            const int i = idx / H, c = idx - i * H;
            const float x1 = X1[idx];
            float d = 0.f;
            for (int j = 0; j < N2; ++j) {
                const float at = s.att[i * N2 + j];
                d = fmaf(at, s.dM1[j * H + c], d);
                if (at * x1 == s.Z1[j * H + c]) d = fmaf(at, s.dZ1[j * H + c], d);
            }
            dX1[idx] += d;
        }
        __syncthreads();
        // B2: att = normalize(X1) normalize(X2)^T
        FOR(i, N1) { float q = 0.f; for (int j = 0; j < N2; ++j) q = fmaf(s.dAtt[i * N2 + j], s.att[i * N2 + j], q); s.T1[i] = q; }
        FOR(j, N2) { float q = 0.f; for (int i = 0; i < N1; ++i) q = fmaf(s.dAtt[i * N2 + j], s.att[i * N2 + j], q); s.T2[j] = q; }
        __syncthreads();
        FOR(idx, (long)N1 * H) {
            const int i = idx / H, c = idx - i * H;
            const float n = s.xn1[i], sc = 1.f / (n + EPS_N);
            float d = 0.f;
            for (int j = 0; j < N2; ++j) d = fmaf(s.dAtt[i * N2 + j], X2[j * H + c] / (s.xn2[j] + EPS_N), d);
            if (n > 0.f) dX1[idx] += sc * d - X1[idx] * s.T1[i] * sc / n;
        }
        FOR(idx, (long)N2 * H) {
            const int j = idx / H, c = idx - j * H;
            const float n = s.xn2[j], sc = 1.f / (n + EPS_N);
            float d = 0.f;
            for (int i = 0; i < N1; ++i) d = fmaf(s.dAtt[i * N2 + j], X1[i * H + c] / (s.xn1[i] + EPS_N), d);
            if (n > 0.f) dX2[idx] += sc * d - X2[idx] * s.T2[j] * sc / n;
        }
        __syncthreads();
        // B4: max-pooling matching, perspective by perspective
        for (int k = 0; k < K; ++k) {
            maxpool_sim(a, s, X1, X2, k);
            const float g1 = G1[k], g2 = G2[k];
            FOR(idx, (long)N1 * N2) {
                const int i = idx / N2, j = idx - i * N2;
                const float v = s.S[idx];
                s.dS[idx] = (v == s.m1[i * K + k] ? g1 : 0.f) + (v == s.m2[j * K + k] ? g2 : 0.f);
            }
            __syncthreads();
            FOR(i, N1) { float q = 0.f; for (int j = 0; j < N2; ++j) q = fmaf(s.dS[i * N2 + j], s.S[i * N2 + j], q); s.T1[i] = q; }
            FOR(j, N2) { float q = 0.f; for (int i = 0; i < N1; ++i) q = fmaf(s.dS[i * N2 + j], s.S[i * N2 + j], q); s.T2[j] = q; }
            __syncthreads();
            const float *w = a.W[0] + (long)k * H;
            float *dWs = sm_dw + (long)k * H;
            FOR(idx, (long)N1 * H) {
                const int i = idx / H, c = idx - i * H;
                const float n = s.nm1[i * K + k], sc = 1.f / (n + EPS_N), x = X1[idx], p = w[c] * x;
                float d = 0.f;
                for (int j = 0; j < N2; ++j) d = fmaf(s.dS[i * N2 + j], w[c] * X2[j * H + c] / (s.nm2[j * K + k] + EPS_N), d);
                const float dp = n > 0.f ? sc * d - p * s.T1[i] * sc / n : 0.f;
                dX1[idx] += w[c] * dp;
                atomicAdd(dWs + c, dp * x);
            }
            FOR(idx, (long)N2 * H) {
                const int j = idx / H, c = idx - j * H;
                const float n = s.nm2[j * K + k], sc = 1.f / (n + EPS_N), x = X2[idx], p = w[c] * x;
                float d = 0.f;
                for (int i = 0; i < N1; ++i) d = fmaf(s.dS[i * N2 + j], w[c] * X1[i * H + c] / (s.nm1[i * K + k] + EPS_N), d);
                const float dp = n > 0.f ? sc * d - p * s.T2[j] * sc / n : 0.f;
                dX2[idx] += w[c] * dp;
                atomicAdd(dWs + c, dp * x);
            }
            __syncthreads();
        }
    }
    if (BWD) {
        __syncthreads();
        for (int t = 0; t < 3; ++t)
            if (a.dW[t]) FOR(idx, (long)K * H) atomicAdd(a.dW[t] + idx, sm_dw[(long)t * K * H + idx]);
    }
}
#undef FOR

}  // namespace bimpm
}  // namespace bmp

using namespace bmp;

static int bimpm_grid(int mb) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int cap = 2 * sms;
    return mb < cap ? mb : cap;
}

extern "C" size_t bmp_bimpm_workspace_bytes(int mb, int n1, int n2, int hidden, int head) {
    if (mb <= 0 || n1 <= 0 || n2 <= 0 || hidden <= 0 || head <= 0) return 0;
    return (size_t)bimpm_grid(mb) * (size_t)bimpm::scratch_floats(n1, n2, hidden, head) * sizeof(float);
}

static int bimpm_check(const bmp_bimpm_t *a) {
    if (!a || !a->atoms_1 || !a->atoms_2 || !a->max_pooling_W || !a->att_mean_W || !a->att_max_W || !a->workspace) {
        set_error("bmp_bimpm: null pointer");
        return BMP_EINVAL;
    }
    if (a->n1 <= 0 || a->n2 <= 0 || a->n1 > BMP_MAX_ATOMS || a->n2 > BMP_MAX_ATOMS || a->hidden <= 0 || a->head <= 0 ||
        (long)a->head * a->hidden > 16384) {
        set_error("bmp_bimpm: shape N1=%d N2=%d hidden=%d head=%d outside (N <= 64, head * hidden <= 16384)", a->n1, a->n2, a->hidden, a->head);
        return BMP_ESHAPE;
    }
    if (a->workspace_bytes < bmp_bimpm_workspace_bytes(a->mb, a->n1, a->n2, a->hidden, a->head)) {
        set_error("bmp_bimpm: workspace of >= %zu bytes required", bmp_bimpm_workspace_bytes(a->mb, a->n1, a->n2, a->hidden, a->head));
        return BMP_EINVAL;
    }
    return BMP_OK;
}

static bimpm::Args bimpm_args(const bmp_bimpm_t *a) {
    bimpm::Args k = {};
    k.mb = a->mb; k.N1 = a->n1; k.N2 = a->n2; k.H = a->hidden; k.K = a->head;
    k.X1 = a->atoms_1; k.X2 = a->atoms_2;
    k.W[0] = a->max_pooling_W; k.W[1] = a->att_mean_W; k.W[2] = a->att_max_W;
    k.scratch = (float *)a->workspace;
    k.scratch_floats = bimpm::scratch_floats(a->n1, a->n2, a->hidden, a->head);
    return k;
}

extern "C" int bmp_bimpm_forward(const bmp_bimpm_t *a, void *stream) {
    int rc = bimpm_check(a);
    if (rc) return rc;
    if (a->mb <= 0) return BMP_OK;
    if (!a->out_1 || !a->out_2) { set_error("bmp_bimpm_forward: null output"); return BMP_EINVAL; }
    bimpm::Args k = bimpm_args(a);
    k.out1 = a->out_1; k.out2 = a->out_2;
    bimpm::bimpm_kernel<false><<<bimpm_grid(a->mb), bimpm::NT, 0, (cudaStream_t)stream>>>(k);
    count_launch();
    return check_launch("bimpm_kernel<fwd>");
}

extern "C" int bmp_bimpm_backward(const bmp_bimpm_t *a, void *stream) {
    int rc = bimpm_check(a);
    if (rc) return rc;
    if (a->mb <= 0) return BMP_OK;
    if (!a->d_out_1 || !a->d_out_2 || !a->d_atoms_1 || !a->d_atoms_2) { set_error("bmp_bimpm_backward: null gradient buffer"); return BMP_EINVAL; }
    bimpm::Args k = bimpm_args(a);
    k.G1 = a->d_out_1; k.G2 = a->d_out_2; k.dX1 = a->d_atoms_1; k.dX2 = a->d_atoms_2;
    k.dW[0] = a->d_max_pooling_W; k.dW[1] = a->d_att_mean_W; k.dW[2] = a->d_att_max_W;
    const int smem = 3 * a->head * a->hidden * (int)sizeof(float);
    cudaFuncSetAttribute(bimpm::bimpm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bimpm::bimpm_kernel<true><<<bimpm_grid(a->mb), bimpm::NT, smem, (cudaStream_t)stream>>>(k);
    count_launch();
    return check_launch("bimpm_kernel<bwd>");
}
