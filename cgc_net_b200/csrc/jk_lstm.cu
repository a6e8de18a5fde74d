// Fused DenseJK: the jumping-knowledge aggregation of one stage (reference model/network.py:11-55).
//
//   per node: x_t = feats[n, t*C:(t+1)*C], t < L;  bi-LSTM(C -> H = C*L/2) over the L "time steps" (PyTorch gate
//   order i, f, g, o; zero initial state); score_t = att_w . [h_fwd_t, h_bwd_t] + att_b; alpha = softmax_t(score);
//   out[n] = sum_t alpha_t x_t.
//
// The reference runs this as a cuDNN LSTM over a [rows, 3, 20] "batch" plus four element-wise kernels; here ONE
// thread owns one node and walks both directions with the recurrent state in registers / thread-private shared
// memory columns; the weights live in shared memory and are read as warp-wide broadcast LDS.128 (every lane needs
// the same weight at the same time).  In training mode the activated gates, cell and hidden states and the
// attention weights go to a tape (node-fastest layout: all tape traffic is coalesced) from which the backward
// kernel produces d feats, the gate pre-activation gradients and the score gradients; weight / bias gradients are
// contractions of those over the nodes and run on the library's GEMM (see ops.DenseJKFunction).
#include "common.cuh"

namespace cgc {

constexpr int kJkThreads = 128;

template <int C, int H, int L>
struct JkShape {
  static constexpr int HP = (H + 3) & ~3;          // hidden width padded to float4
  static constexpr int G4 = 4 * H;                 // gate rows per direction
  static constexpr int kWih = 2 * G4 * C;
  static constexpr int kWhh = 2 * G4 * HP;
  static constexpr int kBias = 2 * G4;
  static constexpr int kAtt = ((2 * H + 1) + 3) & ~3;
  static constexpr int kWeights = kWih + kWhh + kBias + kAtt;
  static constexpr int kPerThread = 2 * H;         // two thread-private columns of H floats
  static constexpr size_t kSmemBytes = sizeof(float) * (size_t)(kWeights + kPerThread * kJkThreads);
};

// weights -> shared memory: W_ih [2][4H][C], W_hh [2][4H][HP] (zero padded), bias = b_ih + b_hh, att_w, att_b
template <int C, int H, int L>
__device__ __forceinline__ void jk_stage_weights(float* s, const float* const* wih, const float* const* whh,
                                                 const float* const* bih, const float* const* bhh,
                                                 const float* att_w, const float* att_b) {
  using S = JkShape<C, H, L>;
  float* s_wih = s;
  float* s_whh = s_wih + S::kWih;
  float* s_b = s_whh + S::kWhh;
  float* s_att = s_b + S::kBias;
  for (int i = threadIdx.x; i < S::kWih; i += kJkThreads) {
    const int d = i / (S::G4 * C), r = i - d * S::G4 * C;
    s_wih[i] = __ldg(wih[d] + r);
  }
  for (int i = threadIdx.x; i < S::kWhh; i += kJkThreads) {
    const int d = i / (S::G4 * S::HP), r = i - d * S::G4 * S::HP;
    const int row = r / S::HP, k = r - row * S::HP;
    s_whh[i] = k < H ? __ldg(whh[d] + row * H + k) : 0.f;
  }
  for (int i = threadIdx.x; i < S::kBias; i += kJkThreads) {
    const int d = i / S::G4, r = i - d * S::G4;
    s_b[i] = (bih ? __ldg(bih[d] + r) : 0.f) + (bhh ? __ldg(bhh[d] + r) : 0.f);
  }
  for (int i = threadIdx.x; i < S::kAtt; i += kJkThreads)
    s_att[i] = i < 2 * H ? __ldg(att_w + i) : (i == 2 * H && att_b ? __ldg(att_b) : 0.f);
}

struct JkWeights {
  const float* wih[2];
  const float* whh[2];
  const float* bih[2];
  const float* bhh[2];
  const float* att_w;
  const float* att_b;
};

__device__ __forceinline__ float sigmoidf_(float a) { return 1.0f / (1.0f + expf(-a)); }

// tape sections (floats): gates [2][L][5][H][rows] (i, f, g, o, c), hs [2][L][H][rows], alpha [L][rows]
template <int C, int H, int L>
__global__ void __launch_bounds__(kJkThreads, 2)
k_jk_fwd(const float* __restrict__ feats, int64_t ld, int64_t rows, JkWeights wts, float* __restrict__ out,
         int64_t ldo, float* __restrict__ tape) {
  using S = JkShape<C, H, L>;
  extern __shared__ __align__(16) float smem[];
  float* s_wih = smem;
  float* s_whh = s_wih + S::kWih;
  float* s_b = s_whh + S::kWhh;
  float* s_att = s_b + S::kBias;
  float* s_hn = s_att + S::kAtt;                       // [H][threads]
  float* s_c = s_hn + H * kJkThreads;                  // [H][threads]
  jk_stage_weights<C, H, L>(smem, wts.wih, wts.whh, wts.bih, wts.bhh, wts.att_w, wts.att_b);
  __syncthreads();
  const int tid = threadIdx.x;
  const int64_t n = (int64_t)blockIdx.x * kJkThreads + tid;
  if (n >= rows) return;
  const float* frow = feats + n * ld;
  float* t_gates = tape;
  float* t_hs = tape ? tape + (int64_t)2 * L * 5 * H * rows : nullptr;
  float* t_alpha = tape ? t_hs + (int64_t)2 * L * H * rows : nullptr;

  float score[L];
#pragma unroll
  for (int t = 0; t < L; ++t) score[t] = s_att[2 * H];
  for (int d = 0; d < 2; ++d) {
    float hp[S::HP];
#pragma unroll
    for (int k = 0; k < S::HP; ++k) hp[k] = 0.f;
    for (int j = 0; j < H; ++j) s_c[j * kJkThreads + tid] = 0.f;
    for (int pos = 0; pos < L; ++pos) {
      const int t = d == 0 ? pos : L - 1 - pos;
      float x[C];
#pragma unroll
      for (int k = 0; k < C; ++k) x[k] = __ldg(frow + t * C + k);
      float sc = 0.f;
#pragma unroll 1
      for (int j = 0; j < H; ++j) {
        float a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = d * S::G4 + q * H + j;
          float acc = s_b[row];
          const float4* w1 = reinterpret_cast<const float4*>(s_wih + row * C);
#pragma unroll
          for (int k4 = 0; k4 < C / 4; ++k4) {
            const float4 w = w1[k4];
            acc = fmaf(w.x, x[4 * k4], acc); acc = fmaf(w.y, x[4 * k4 + 1], acc);
            acc = fmaf(w.z, x[4 * k4 + 2], acc); acc = fmaf(w.w, x[4 * k4 + 3], acc);
          }
          const float4* w2 = reinterpret_cast<const float4*>(s_whh + row * S::HP);
#pragma unroll
          for (int k4 = 0; k4 < S::HP / 4; ++k4) {
            const float4 w = w2[k4];
            acc = fmaf(w.x, hp[4 * k4], acc); acc = fmaf(w.y, hp[4 * k4 + 1], acc);
            acc = fmaf(w.z, hp[4 * k4 + 2], acc); acc = fmaf(w.w, hp[4 * k4 + 3], acc);
          }
          a[q] = acc;
        }
        const float gi = sigmoidf_(a[0]), gf = sigmoidf_(a[1]), gg = tanhf(a[2]), go = sigmoidf_(a[3]);
        const float cn = fmaf(gf, s_c[j * kJkThreads + tid], gi * gg);
        const float hn = go * tanhf(cn);
        s_c[j * kJkThreads + tid] = cn;
        s_hn[j * kJkThreads + tid] = hn;
        sc = fmaf(s_att[d * H + j], hn, sc);
        if (tape) {
          float* tg = t_gates + ((((int64_t)d * L + t) * 5) * H + j) * rows + n;
          tg[0] = gi;
          tg[(int64_t)H * rows] = gf;
          tg[(int64_t)2 * H * rows] = gg;
          tg[(int64_t)3 * H * rows] = go;
          tg[(int64_t)4 * H * rows] = cn;
          t_hs[(((int64_t)d * L + t) * H + j) * rows + n] = hn;
        }
      }
#pragma unroll
      for (int k = 0; k < H; ++k) hp[k] = s_hn[k * kJkThreads + tid];
#pragma unroll
      for (int tt = 0; tt < L; ++tt) score[tt] += (tt == t) ? sc : 0.f;
    }
  }
  float mx = score[0];
#pragma unroll
  for (int t = 1; t < L; ++t) mx = fmaxf(mx, score[t]);
  float den = 0.f, alpha[L];
#pragma unroll
  for (int t = 0; t < L; ++t) { alpha[t] = expf(score[t] - mx); den += alpha[t]; }
  const float inv = 1.0f / den;
#pragma unroll
  for (int t = 0; t < L; ++t) {
    alpha[t] *= inv;
    if (tape) t_alpha[(int64_t)t * rows + n] = alpha[t];
  }
  float* orow = out + n * ldo;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < L; ++t) v = fmaf(alpha[t], __ldg(frow + t * C + k), v);
    orow[k] = v;
  }
}

template <int C, int H, int L>
__global__ void __launch_bounds__(kJkThreads, 2)
k_jk_bwd(const float* __restrict__ g_out, int64_t ldg, const float* __restrict__ feats, int64_t ld, int64_t rows,
         JkWeights wts, const float* __restrict__ tape, float* __restrict__ g_feats, int64_t ldgf,
         float* __restrict__ dgates, float* __restrict__ dscore) {
  using S = JkShape<C, H, L>;
  extern __shared__ __align__(16) float smem[];
  float* s_wih = smem;
  float* s_whh = s_wih + S::kWih;
  float* s_b = s_whh + S::kWhh;
  float* s_att = s_b + S::kBias;
  float* s_dh = s_att + S::kAtt;                       // [H][threads]  d loss / d h of the step being undone
  float* s_dc = s_dh + H * kJkThreads;                 // [H][threads]  d loss / d c carried backwards
  jk_stage_weights<C, H, L>(smem, wts.wih, wts.whh, nullptr, nullptr, wts.att_w, nullptr);
  __syncthreads();
  const int tid = threadIdx.x;
  const int64_t n = (int64_t)blockIdx.x * kJkThreads + tid;
  if (n >= rows) return;
  const float* frow = feats + n * ld;
  const float* t_gates = tape;
  const float* t_alpha = tape + (int64_t)2 * L * 5 * H * rows + (int64_t)2 * L * H * rows;

  // attention: d alpha_t = g . x_t;  d score = alpha * (d alpha - sum alpha d alpha);  gx_t = alpha_t g
  float gx[L][C];
  float ds[L];
  {
    float g[C], alpha[L], dal[L];
#pragma unroll
    for (int k = 0; k < C; ++k) g[k] = __ldg(g_out + n * ldg + k);
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < L; ++t) {
      alpha[t] = __ldg(t_alpha + (int64_t)t * rows + n);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < C; ++k) s = fmaf(g[k], __ldg(frow + t * C + k), s);
      dal[t] = s;
      dot = fmaf(alpha[t], s, dot);
    }
#pragma unroll
    for (int t = 0; t < L; ++t) {
      ds[t] = alpha[t] * (dal[t] - dot);
      dscore[(int64_t)t * rows + n] = ds[t];
#pragma unroll
      for (int k = 0; k < C; ++k) gx[t][k] = alpha[t] * g[k];
    }
  }
  for (int d = 0; d < 2; ++d) {
    for (int j = 0; j < H; ++j) {
      s_dh[j * kJkThreads + tid] = 0.f;
      s_dc[j * kJkThreads + tid] = 0.f;
    }
    for (int pos = L - 1; pos >= 0; --pos) {
      const int t = d == 0 ? pos : L - 1 - pos;
      const int tprev = d == 0 ? t - 1 : t + 1;        // the step computed before this one (pos > 0)
      float dst = 0.f;
#pragma unroll
      for (int tt = 0; tt < L; ++tt) dst += (tt == t) ? ds[tt] : 0.f;
      float dx[C], dhp[S::HP];
#pragma unroll
      for (int k = 0; k < C; ++k) dx[k] = 0.f;
#pragma unroll
      for (int k = 0; k < S::HP; ++k) dhp[k] = 0.f;
#pragma unroll 1
      for (int j = 0; j < H; ++j) {
        const float* tg = t_gates + ((((int64_t)d * L + t) * 5) * H + j) * rows + n;
        const float gi = __ldg(tg), gf = __ldg(tg + (int64_t)H * rows), gg = __ldg(tg + (int64_t)2 * H * rows);
        const float go = __ldg(tg + (int64_t)3 * H * rows), cn = __ldg(tg + (int64_t)4 * H * rows);
        const float cprev =
            pos > 0 ? __ldg(t_gates + ((((int64_t)d * L + tprev) * 5 + 4) * H + j) * rows + n) : 0.f;
        const float dh = fmaf(dst, s_att[d * H + j], s_dh[j * kJkThreads + tid]);
        const float tc = tanhf(cn);
        const float dc = fmaf(dh * go, 1.0f - tc * tc, s_dc[j * kJkThreads + tid]);
        float da[4];
        da[0] = dc * gg * gi * (1.0f - gi);
        da[1] = dc * cprev * gf * (1.0f - gf);
        da[2] = dc * gi * (1.0f - gg * gg);
        da[3] = dh * tc * go * (1.0f - go);
        s_dc[j * kJkThreads + tid] = dc * gf;
        float* dgp = dgates + (((int64_t)d * L + t) * S::G4 + j) * rows + n;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dgp[(int64_t)q * H * rows] = da[q];
          const int row = d * S::G4 + q * H + j;
          const float4* w1 = reinterpret_cast<const float4*>(s_wih + row * C);
#pragma unroll
          for (int k4 = 0; k4 < C / 4; ++k4) {
            const float4 w = w1[k4];
            dx[4 * k4] = fmaf(w.x, da[q], dx[4 * k4]); dx[4 * k4 + 1] = fmaf(w.y, da[q], dx[4 * k4 + 1]);
            dx[4 * k4 + 2] = fmaf(w.z, da[q], dx[4 * k4 + 2]); dx[4 * k4 + 3] = fmaf(w.w, da[q], dx[4 * k4 + 3]);
          }
          const float4* w2 = reinterpret_cast<const float4*>(s_whh + row * S::HP);
#pragma unroll
          for (int k4 = 0; k4 < S::HP / 4; ++k4) {
            const float4 w = w2[k4];
            dhp[4 * k4] = fmaf(w.x, da[q], dhp[4 * k4]); dhp[4 * k4 + 1] = fmaf(w.y, da[q], dhp[4 * k4 + 1]);
            dhp[4 * k4 + 2] = fmaf(w.z, da[q], dhp[4 * k4 + 2]);
            dhp[4 * k4 + 3] = fmaf(w.w, da[q], dhp[4 * k4 + 3]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < H; ++k) s_dh[k * kJkThreads + tid] = dhp[k];
#pragma unroll
      for (int tt = 0; tt < L; ++tt)
#pragma unroll
        for (int k = 0; k < C; ++k) gx[tt][k] += (tt == t) ? dx[k] : 0.f;
    }
  }
  float* grow = g_feats + n * ldgf;
#pragma unroll
  for (int t = 0; t < L; ++t)
#pragma unroll
    for (int k = 0; k < C; ++k) grow[t * C + k] = gx[t][k];
}

}  // namespace cgc

using namespace cgc;

// Instantiated shapes: `--hidden-dim` values (train.py:393) whose LSTM weights fit in shared memory next to the
// thread-private state columns: C = 16 / 20 / 32 (hidden 24 / 30 / 48), 3 layers.  C = 64 would need 491 KB of
// weights (W_ih 2 x 384 x 64, W_hh 2 x 384 x 96) and stays on the library LSTM.
static bool jk_supported(int channels, int layers) {
  return layers == 3 && (channels == 16 || channels == 20 || channels == 32);
}
#define CGC_JK_DISPATCH(channels, CALL)       \
  switch (channels) {                         \
    case 16: { CALL(16, 24) } break;          \
    case 20: { CALL(20, 30) } break;          \
    default: { CALL(32, 48) } break;          \
  }

extern "C" int cgc_jk_supported(int32_t channels, int32_t layers) { return jk_supported(channels, layers) ? 1 : 0; }

extern "C" int64_t cgc_jk_tape_floats(int64_t rows, int32_t channels, int32_t layers) {
  const int64_t hid = (int64_t)channels * layers / 2;
  return rows * (2 * layers * 5 * hid + 2 * layers * hid + layers);
}

static JkWeights jk_weights(const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                            const float* w_ih_b, const float* w_hh_b, const float* b_ih_b, const float* b_hh_b,
                            const float* att_w, const float* att_b) {
  JkWeights w;
  w.wih[0] = w_ih_f; w.wih[1] = w_ih_b;
  w.whh[0] = w_hh_f; w.whh[1] = w_hh_b;
  w.bih[0] = b_ih_f; w.bih[1] = b_ih_b;
  w.bhh[0] = b_hh_f; w.bhh[1] = b_hh_b;
  w.att_w = att_w; w.att_b = att_b;
  return w;
}

extern "C" int cgc_jk_lstm_fwd(const float* feats, int64_t ld, int64_t rows, int32_t channels, int32_t layers,
                               const float* w_ih_f, const float* w_hh_f, const float* b_ih_f, const float* b_hh_f,
                               const float* w_ih_b, const float* w_hh_b, const float* b_ih_b, const float* b_hh_b,
                               const float* att_w, const float* att_b, float* out, int64_t ldo, float* tape,
                               void* stream) {
  CGC_REQUIRE(feats && w_ih_f && w_hh_f && b_ih_f && b_hh_f && w_ih_b && w_hh_b && b_ih_b && b_hh_b && att_w &&
                  att_b && out && rows > 0,
              CGC_ERR_ARG, "cgc_jk_lstm_fwd: bad args");
  CGC_REQUIRE(jk_supported(channels, layers), CGC_ERR_UNSUPPORTED,
              "cgc_jk_lstm_fwd: instantiated for channels 16 / 20 / 32, layers=3 (got %d, %d)", channels, layers);
  CGC_REQUIRE(ld >= (int64_t)channels * layers && ldo >= channels, CGC_ERR_ARG, "cgc_jk_lstm_fwd: leading dims");
  const JkWeights wts = jk_weights(w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_b, w_hh_b, b_ih_b, b_hh_b, att_w, att_b);
#define CGC_JK_FWD(C_, H_)                                                                                         \
  using S = JkShape<C_, H_, 3>;                                                                                    \
  auto kern = k_jk_fwd<C_, H_, 3>;                                                                                 \
  CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kSmemBytes));          \
  kern<<<(unsigned)ceil_div(rows, kJkThreads), kJkThreads, S::kSmemBytes, as_stream(stream)>>>(feats, ld, rows, wts, \
                                                                                               out, ldo, tape);
  CGC_JK_DISPATCH(channels, CGC_JK_FWD)
#undef CGC_JK_FWD
  CGC_CHECK_LAUNCH("k_jk_fwd");
  return 0;
}

extern "C" int cgc_jk_lstm_bwd(const float* g_out, int64_t ldg, const float* feats, int64_t ld, int64_t rows,
                               int32_t channels, int32_t layers, const float* w_ih_f, const float* w_hh_f,
                               const float* w_ih_b, const float* w_hh_b, const float* att_w, const float* tape,
                               float* g_feats, int64_t ldgf, float* dgates, float* dscore, void* stream) {
  CGC_REQUIRE(g_out && feats && w_ih_f && w_hh_f && w_ih_b && w_hh_b && att_w && tape && g_feats && dgates &&
                  dscore && rows > 0,
              CGC_ERR_ARG, "cgc_jk_lstm_bwd: bad args");
  CGC_REQUIRE(jk_supported(channels, layers), CGC_ERR_UNSUPPORTED,
              "cgc_jk_lstm_bwd: instantiated for channels 16 / 20 / 32, layers=3 (got %d, %d)", channels, layers);
  CGC_REQUIRE(ld >= (int64_t)channels * layers && ldgf >= (int64_t)channels * layers && ldg >= channels, CGC_ERR_ARG,
              "cgc_jk_lstm_bwd: leading dims");
  const JkWeights wts = jk_weights(w_ih_f, w_hh_f, nullptr, nullptr, w_ih_b, w_hh_b, nullptr, nullptr, att_w, nullptr);
#define CGC_JK_BWD(C_, H_)                                                                                         \
  using S = JkShape<C_, H_, 3>;                                                                                    \
  auto kern = k_jk_bwd<C_, H_, 3>;                                                                                 \
  CGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kSmemBytes));          \
  kern<<<(unsigned)ceil_div(rows, kJkThreads), kJkThreads, S::kSmemBytes, as_stream(stream)>>>(                    \
      g_out, ldg, feats, ld, rows, wts, tape, g_feats, ldgf, dgates, dscore);
  CGC_JK_DISPATCH(channels, CGC_JK_BWD)
#undef CGC_JK_BWD
  CGC_CHECK_LAUNCH("k_jk_bwd");
  return 0;
}
