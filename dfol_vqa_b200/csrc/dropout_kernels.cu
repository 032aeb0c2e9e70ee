// Training-mode dropout of the visual oracle's networks (nn.Dropout in front of every Linear of RegularMLP /
// EmbeddingLayer: nsvqa/nn/vision/regular_mlp.py:29-32, embedding_layer.py:73; sample_config.yaml: dropout 0.1).
//
// The keep/drop decision of element (row, col) of dropout site `site` is a pure function of (seed, site, row, col):
// a counter-based hash keyed by the seed, counter = (group of 8 consecutive columns of the row, site), 16 bits per element,
// keep iff bits >= round(p * 65536); kept elements are scaled by 1 / (1 - p) like torch.nn.functional.dropout.
// Nothing is stored: every kernel that needs a mask recomputes it, and the parity tests export the very same masks
// (dfol_dropout_scale on a tensor of ones) for the CPU oracle -- the reference's own RNG stream cannot be matched.
#include <cuda_bf16.h>

#include "dfol_common.cuh"

namespace dfol {

// Four 32-bit words (= 8 elements of 16 bits) per counter: the counter, the site and the seed are folded into one
// word with odd multipliers, and each output word is a full-avalanche finalizer (the 32-bit finalizer of MurmurHash3:
// xor-shift 16, multiply, xor-shift 13, multiply, xor-shift 16) of that word advanced by a distinct odd constant and
// keyed by the second seed word.  ~40 integer instructions per 8 elements; the Philox4x32-10 of round 1 (~80) made
// the mask generation the issue-bound part of dfol_pair_features_dropout (642 M elements per c1 batch).  A dropout
// mask needs independent-looking Bernoulli(1 - p) decisions, not a cryptographic stream; tests/test_gpu_dropout.py
// checks the keep rate per site, row and column and the independence of neighbouring elements and sites.
__device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint4 drop_words(uint32_t g_lo, uint32_t g_hi, uint32_t site, uint2 k) {
  const uint32_t h = fmix32(g_lo * 0x9E3779B1u + (g_hi * 0x85EBCA77u ^ site * 0xC2B2AE3Du ^ k.x));
  return make_uint4(fmix32(h ^ k.y), fmix32((h + 0x27D4EB2Fu) ^ k.y), fmix32((h + 0x165667B1u) ^ k.y),
                    fmix32((h + 0x7F4A7C15u) ^ k.y));
}

struct DropSite {
  uint2 key;
  uint32_t site, thresh;
  long long groups_per_row;
  float scale;
};

// scale factors (0 or 1/(1-p)) of the 8 elements of column group `g8` of `row`
__device__ __forceinline__ void drop_scales8(const DropSite& d, long long row, long long g8, float s[8]) {
  const unsigned long long g = (unsigned long long)row * (unsigned long long)d.groups_per_row + (unsigned long long)g8;
  const uint4 r = drop_words((uint32_t)g, (uint32_t)(g >> 32), d.site, d.key);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[2 * i] = ((w[i] & 0xffffu) >= d.thresh) ? d.scale : 0.0f;
    s[2 * i + 1] = ((w[i] >> 16) >= d.thresh) ? d.scale : 0.0f;
  }
}

static DropSite make_site(unsigned long long seed, int site, float p, long long cols) {
  DropSite d;
  d.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  d.site = (uint32_t)site;
  long long t = (long long)(p * 65536.0 + 0.5);
  d.thresh = (uint32_t)(t < 0 ? 0 : (t > 65536 ? 65536 : t));
  d.groups_per_row = (cols + 7) / 8;
  d.scale = 1.0f / (1.0f - p);
  return d;
}

// x[row, col] *= scale(row, col), in place; one thread per group of 8 columns
template <class T>
__global__ void __launch_bounds__(256) dropout_scale_kernel(T* __restrict__ x, long long ld, long long rows, int cols,
                                                            DropSite d) {
  const long long total = rows * d.groups_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / d.groups_per_row, g8 = i - row * d.groups_per_row;
    float s[8];
    drop_scales8(d, row, g8, s);
    T* p = x + row * ld + g8 * 8;
    const int m = min(8, cols - (int)(g8 * 8));
    if (sizeof(T) == 2 && m == 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      uint4 v = *reinterpret_cast<uint4*>(p);
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        h[j] = __floats2bfloat162_rn(f.x * s[2 * j], f.y * s[2 * j + 1]);
      }
      *reinterpret_cast<uint4*>(p) = v;
    } else {
      for (int j = 0; j < m; ++j) p[j] = (T)((float)p[j] * s[j]);
    }
  }
}

// geometry features of an ordered pair (s,o): dist, asin(dy/dist), sign(x_o-x_s), sign(y_o-y_s)
// (batch_gqa_boxfeatures_pipeline.py:260-279; same arithmetic as scene_kernels.cu pair_geometry)
__device__ __forceinline__ void drop_pair_geometry(const float* ps, const float* po, float g[4]) {
  const float x1 = ps[0], y1 = ps[1], w1 = ps[2], h1 = ps[3];
  const float x2 = po[0], y2 = po[1], w2 = po[2], h2 = po[3];
  const float dx = x1 + w1 / 2.0f - x2 - w2 / 2.0f;
  const float dy = y1 + h1 / 2.0f - y2 - h2 / 2.0f;
  const float dist = sqrtf(dx * dx + dy * dy);
  g[0] = dist;
  g[1] = asinf(dy / fmaxf(dist, 1e-10f));
  const float sx = x2 - x1, sy = y2 - y1;
  g[2] = (sx > 0.0f) ? 1.0f : (sx < 0.0f ? -1.0f : 0.0f);
  g[3] = (sy > 0.0f) ? 1.0f : (sy < 0.0f ? -1.0f : 0.0f);
}

// Masked relation-network input rows: out[(b,s,o), :] = mask .* [obj_s | obj_o | geo(s,o)], zero beyond 2*ldo+4.
// With an independent mask per pair element the first layer no longer factors into U[s] + V[o], so the training-mode
// dropout path evaluates it as the reference does, on the materialised pair matrix.  One warp per pair row.
template <class T>
__global__ void __launch_bounds__(256) pair_features_dropout_kernel(
    const float* __restrict__ obj, long long ldobj, int width, int pos_col, T* __restrict__ out, long long ldout,
    int out_cols, const int32_t* __restrict__ pair_row, const int32_t* __restrict__ obj_row,
    const int32_t* __restrict__ img_n, const int32_t* __restrict__ pair_img, long long pairs, DropSite d) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int in_cols = 2 * width + 4;
  const bool vec4 = (width % 4) == 0 && (ldobj % 4) == 0 && (reinterpret_cast<uintptr_t>(obj) & 15) == 0;
  for (long long r = warp; r < pairs; r += nwarps) {
    const int b = pair_img[r];
    const int n = img_n[b];
    const int local = (int)(r - pair_row[b]);
    const int s = local / n, o = local - s * n;
    const float* os = obj + (long long)(obj_row[b] + s) * ldobj;
    const float* oo = obj + (long long)(obj_row[b] + o) * ldobj;
    float geo[4];
    drop_pair_geometry(os + pos_col, oo + pos_col, geo);
    T* dst = out + r * ldout;
    for (int g8 = lane; g8 * 8 < out_cols; g8 += 32) {
      float sc[8];
      drop_scales8(d, r, g8, sc);
      float v[8];
      if (vec4) {
        // width % 4 == 0 and 16-byte aligned object rows: each half of the group lies inside ONE segment
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          const int c = g8 * 8 + 4 * hlf;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < width) x = *reinterpret_cast<const float4*>(os + c);
          else if (c < 2 * width) x = *reinterpret_cast<const float4*>(oo + (c - width));
          else if (c < in_cols) x = make_float4(geo[0], geo[1], geo[2], geo[3]);
          v[4 * hlf] = x.x * sc[4 * hlf]; v[4 * hlf + 1] = x.y * sc[4 * hlf + 1];
          v[4 * hlf + 2] = x.z * sc[4 * hlf + 2]; v[4 * hlf + 3] = x.w * sc[4 * hlf + 3];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = g8 * 8 + j;
          float x = 0.0f;
          if (c < width) x = os[c];
          else if (c < 2 * width) x = oo[c - width];
          else if (c < in_cols) x = geo[c - 2 * width];
          v[j] = x * sc[j];
        }
      }
      T* p8 = dst + g8 * 8;
      if (sizeof(T) == 2 && g8 * 8 + 8 <= out_cols && (reinterpret_cast<uintptr_t>(p8) & 15) == 0) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(p8) = pk;
      } else {
        for (int j = 0; j < 8 && g8 * 8 + j < out_cols; ++j) p8[j] = (T)v[j];
      }
    }
  }
}

// The same for the tensor-core path (bf16 rows, width % 4 == 0, 16-byte aligned rows), restructured around what repeats:
// a block owns PFD_ROWS consecutive pair rows, a thread owns ONE column group of 8 and every second row, so that
//   * the subject half of a row ([obj_s], the same for the n rows of a subject) stays in registers,
//   * the geometry (sqrt, asin) is evaluated by the one thread whose group holds it, not by all 32 lanes of a row warp,
//   * no lanes idle on the ragged tail of 136 groups over 32 lanes (the row-per-warp kernel: 4.25 passes per row),
// and the keep decisions are taken two at a time (packed 16-bit compare) on the bf16 words.  Same masks, same values.
constexpr int PFD_ROWS = 64;

__global__ void __launch_bounds__(288) pair_features_dropout_bf16_kernel(
    const float* __restrict__ obj, long long ldobj, int width, int pos_col, __nv_bfloat16* __restrict__ out,
    long long ldout, int groups, const int32_t* __restrict__ pair_row, const int32_t* __restrict__ obj_row,
    const int32_t* __restrict__ img_n, const int32_t* __restrict__ pair_img, long long pairs, DropSite d) {
  const int rsub = threadIdx.x / groups, g8 = threadIdx.x - rsub * groups;
  if (rsub >= 2) return;
  const long long r_end = min(pairs, ((long long)blockIdx.x + 1) * PFD_ROWS);
  const int in_cols = 2 * width + 4;
  // segment of each half (4 columns) of this thread's group: 0 subject, 1 object, 2 geometry, 3 zero padding
  int seg[2], off[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = g8 * 8 + 4 * h;
    seg[h] = c < width ? 0 : (c < 2 * width ? 1 : (c < in_cols ? 2 : 3));
    off[h] = seg[h] == 0 ? c : c - width;
  }
  // kept elements are scaled by 1 / (1 - p); the drop threshold replicated into both 16-bit halves
  const uint32_t thr2 = d.thresh | (d.thresh << 16);
  int b = -1, n = 1, s = 0, o = 0;
  const float* os = obj;
  const float* ob = obj;
  float4 subj[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  for (long long r = (long long)blockIdx.x * PFD_ROWS + rsub; r < r_end; r += 2) {
    const int bn = pair_img[r];
    bool new_subject = false;
    if (bn != b) {
      b = bn;
      n = img_n[b];
      const int local = (int)(r - pair_row[b]);
      s = local / n;
      o = local - s * n;
      ob = obj + (long long)obj_row[b] * ldobj;
      new_subject = true;
    } else {
      o += 2;
      while (o >= n) { o -= n; ++s; new_subject = true; }
    }
    if (new_subject) {
      os = ob + (long long)s * ldobj;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (seg[h] == 0) subj[h] = *reinterpret_cast<const float4*>(os + off[h]);
    }
    const float* oo = ob + (long long)o * ldobj;
    const unsigned long long g = (unsigned long long)r * (unsigned long long)d.groups_per_row + (unsigned long long)g8;
    const uint4 w = drop_words((uint32_t)g, (uint32_t)(g >> 32), d.site, d.key);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
    uint32_t pk[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 x = subj[h];
      if (seg[h] == 1) x = *reinterpret_cast<const float4*>(oo + off[h]);
      else if (seg[h] == 2) {
        float geo[4];
        drop_pair_geometry(os + pos_col, oo + pos_col, geo);
        x = make_float4(geo[0], geo[1], geo[2], geo[3]);
      } else if (seg[h] == 3) x = make_float4(0.f, 0.f, 0.f, 0.f);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x * d.scale, x.y * d.scale);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(x.z * d.scale, x.w * d.scale);
      // element 2i of the group takes the low, 2i + 1 the high 16 bits of word i; keep iff bits >= threshold
      pk[2 * h] = *reinterpret_cast<const uint32_t*>(&lo) & __vcmpgeu2(ws[2 * h], thr2);
      pk[2 * h + 1] = *reinterpret_cast<const uint32_t*>(&hi) & __vcmpgeu2(ws[2 * h + 1], thr2);
    }
    *reinterpret_cast<uint4*>(out + r * ldout + g8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// Backward of the pair-feature gather: d_obj[t, c] += sum_o dpm[(t,o), c] + sum_s dpm[(s,t), width + c] (+ addend[t, c])
// (the geometry columns carry no parameter gradient).  One block per object row, threads over columns.
template <class T>
__global__ void __launch_bounds__(256) pair_features_bwd_kernel(const T* __restrict__ dpm, long long ld, int width,
                                                                float* __restrict__ d_obj, long long ldobj,
                                                                const __nv_bfloat16* __restrict__ addend,
                                                                long long ld_add, const int32_t* __restrict__ pair_row,
                                                                const int32_t* __restrict__ obj_row,
                                                                const int32_t* __restrict__ img_n,
                                                                const int32_t* __restrict__ obj_img) {
  const long long t = blockIdx.x;
  const int b = obj_img[t];
  const int n = img_n[b];
  const int i = (int)(t - obj_row[b]);
  const T* base = dpm + (long long)pair_row[b] * ld;
  for (int c = threadIdx.x; c < width; c += blockDim.x) {
    float acc = addend ? __bfloat162float(addend[t * ld_add + c]) : 0.f;
    for (int o = 0; o < n; ++o) acc += (float)base[(long long)(i * n + o) * ld + c];
    for (int s2 = 0; s2 < n; ++s2) acc += (float)base[(long long)(s2 * n + i) * ld + width + c];
    d_obj[t * ldobj + c] += acc;
  }
}

}  // namespace dfol

using namespace dfol;

extern "C" int dfol_pair_features_bwd(const void* dpm, int64_t ld, int is_bf16, int width, float* d_obj, int64_t ldobj,
                                      const void* addend_bf16, int64_t ld_add, const int32_t* pair_row,
                                      const int32_t* obj_row, const int32_t* img_n, const int32_t* obj_img,
                                      int64_t objects, void* stream) {
  DFOL_REQUIRE(dpm && d_obj && pair_row && obj_row && img_n && obj_img, "dfol_pair_features_bwd: null pointer");
  DFOL_REQUIRE(width >= 1 && ld >= 2 * width && ldobj >= width && (addend_bf16 == nullptr || ld_add >= width),
               "dfol_pair_features_bwd: bad shape");
  if (objects == 0) return 0;
  const __nv_bfloat16* add = reinterpret_cast<const __nv_bfloat16*>(addend_bf16);
  if (is_bf16)
    pair_features_bwd_kernel<__nv_bfloat16><<<(unsigned)objects, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(dpm), ld, width, d_obj, ldobj, add, ld_add, pair_row, obj_row, img_n,
        obj_img);
  else
    pair_features_bwd_kernel<float><<<(unsigned)objects, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float*>(dpm), ld, width, d_obj, ldobj, add, ld_add, pair_row, obj_row, img_n, obj_img);
  return finish_launch("dfol_pair_features_bwd");
}

extern "C" int dfol_dropout_scale(void* x, int64_t ld, int64_t rows, int cols, int is_bf16, uint64_t seed, int site,
                                  float p, void* stream) {
  DFOL_REQUIRE(x != nullptr && rows >= 0 && cols > 0 && ld >= cols, "dfol_dropout_scale: bad arguments");
  DFOL_REQUIRE(p >= 0.0f && p < 1.0f, "dfol_dropout_scale: p must be in [0, 1)");
  if (rows == 0) return 0;
  const DropSite d = make_site(seed, site, p, cols);
  const long long total = rows * d.groups_per_row;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (is_bf16)
    dropout_scale_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, ld, rows, cols, d);
  else
    dropout_scale_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)x, ld, rows, cols, d);
  return finish_launch("dfol_dropout_scale");
}

extern "C" int dfol_pair_features_dropout(const float* obj, int64_t ldobj, int width, int pos_col, void* out,
                                          int64_t ldout, int out_cols, int is_bf16, const int32_t* pair_row,
                                          const int32_t* obj_row, const int32_t* img_n, const int32_t* pair_img,
                                          int64_t pairs, uint64_t seed, int site, float p, void* stream) {
  DFOL_REQUIRE(obj && out && pair_row && obj_row && img_n && pair_img, "dfol_pair_features_dropout: null pointer");
  DFOL_REQUIRE(width >= 4 && pos_col + 4 <= width && out_cols >= 2 * width + 4 && ldout >= out_cols,
               "dfol_pair_features_dropout: bad shape");
  DFOL_REQUIRE(p >= 0.0f && p < 1.0f, "dfol_pair_features_dropout: p must be in [0, 1)");
  if (pairs == 0) return 0;
  const DropSite d = make_site(seed, site, p, 2 * width + 4);
  const int groups = out_cols / 8;
  if (is_bf16 && (out_cols % 8) == 0 && (ldout % 8) == 0 && (width % 4) == 0 && (ldobj % 4) == 0 && (pos_col % 4) == 0 &&
      2 * groups <= 288 && d.thresh < 65536u && (reinterpret_cast<uintptr_t>(obj) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const long long nblk = (pairs + PFD_ROWS - 1) / PFD_ROWS;
    pair_features_dropout_bf16_kernel<<<(unsigned)nblk, 288, 0, (cudaStream_t)stream>>>(
        obj, ldobj, width, pos_col, (__nv_bfloat16*)out, ldout, groups, pair_row, obj_row, img_n, pair_img, pairs, d);
    return finish_launch("dfol_pair_features_dropout");
  }
  const long long warps = pairs;
  const int blocks = (int)((warps + 7) / 8 < 148 * 16 ? (warps + 7) / 8 : 148 * 16);
  if (is_bf16)
    pair_features_dropout_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        obj, ldobj, width, pos_col, (__nv_bfloat16*)out, ldout, out_cols, pair_row, obj_row, img_n, pair_img, pairs, d);
  else
    pair_features_dropout_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        obj, ldobj, width, pos_col, (float*)out, ldout, out_cols, pair_row, obj_row, img_n, pair_img, pairs, d);
  return finish_launch("dfol_pair_features_dropout");
}
