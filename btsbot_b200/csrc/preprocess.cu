// K1 -- array preparation kernels (HBM-bound, one CTA per alert, whole triplet staged in shared memory).
//   crop_norm : alert_utils.crop_triplets/crop_norm_cutout (alert_utils.py:54-107) + astype(float32) +
//               transpose(0,3,1,2) (inference_example.py:62-64, train.py:139-155, val.py:92-94)
//   pad_norm  : numeric tail of alert_utils.make_triplet (alert_utils.py:147-193)
#include <float.h>

#include "common.cuh"

namespace btsb {

constexpr int kImg = 63;
constexpr int kPix = kImg * kImg;      // 3969
constexpr int kTrip = kPix * 3;        // 11907
constexpr int kPreThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kPreThreads)
crop_norm_kernel(const T* __restrict__ in, int s, int margin, int normalize, int out_hwc, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  __shared__ double red[3][kPreThreads / 32];
  __shared__ double nrm_s[3];
  const int tid = threadIdx.x;
  const int64_t a = blockIdx.x;
  const T* src = in + a * (int64_t)kTrip;

  // coalesced streaming load of the whole HWC triplet
  for (int i = tid; i < kTrip; i += kPreThreads) tile[i] = __ldg(src + i);
  __syncthreads();

  if (normalize) {
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int p = tid; p < s * s; p += kPreThreads) {
      const int y = p / s, x = p - y * s;
      const T* px = tile + ((y + margin) * kImg + (x + margin)) * 3;
      const double v0 = (double)px[0], v1 = (double)px[1], v2 = (double)px[2];
      acc0 += v0 * v0; acc1 += v1 * v1; acc2 += v2 * v2;
    }
    acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2);
    if ((tid & 31) == 0) { red[0][tid >> 5] = acc0; red[1][tid >> 5] = acc1; red[2][tid >> 5] = acc2; }
    __syncthreads();
    if (tid < 3) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kPreThreads / 32; ++w) t += red[tid][w];
      nrm_s[tid] = t;   // sum of squares
    }
    __syncthreads();
  }

  const int ss = s * s;
  float* dst = out + a * (int64_t)(3 * ss);
  double nd[3] = {1.0, 1.0, 1.0};
  if (normalize) {
#pragma unroll
    for (int c = 0; c < 3; ++c) nd[c] = sqrt(nrm_s[c]);
  }
  // i walks the OUTPUT linearly (coalesced stores); NCHW: i = c*ss + p, NHWC: i = p*3 + c
  for (int i = tid; i < 3 * ss; i += kPreThreads) {
    int c, p;
    if (out_hwc) { p = i / 3; c = i - p * 3; } else { c = i / ss; p = i - c * ss; }
    const int y = p / s, x = p - y * s;
    const T v = tile[((y + margin) * kImg + (x + margin)) * 3 + c];
    const double ndc = c == 0 ? nd[0] : (c == 1 ? nd[1] : nd[2]);
    // x / ||x||_2 from the float64 sum of squares, rounded once: within 1 ulp(fp32) of the exact quotient for either
    // input dtype (the reference's float32 path -- numpy's float32 BLAS dot + sqrt, alert_utils.py:75-76 -- has no
    // defined summation order and is itself up to ~1.5 ulp away from it)
    dst[i] = normalize ? (float)((double)v / ndc) : (float)v;
  }
}

// Fast path of K1 for what every caller of the scoring path does (inference_example.py:62-64, train.py:139-155,
// val.py:92-94): astype(float32) + transpose(0,3,1,2) of full 63x63 triplets, no crop, no normalisation.  One CTA per
// alert; the HWC triplet comes in with 16-byte loads (the alert's 11 907 elements start at any 4-/8-byte phase of a
// 16-byte line, so a short scalar head / tail brackets the aligned body and the shared-memory tile is shifted by the
// same phase, which keeps the vector stores into it aligned), and leaves as coalesced rows of the three channel planes
// with compile-time index arithmetic (the generic kernel divides by the runtime crop size per element: it ran at
// 2.6 TB/s, 0.79 of eager PyTorch's permute + contiguous; profiles/r02c).
template <typename T>
__global__ void __launch_bounds__(kPreThreads)
cast_transpose63_kernel(const T* __restrict__ in, float* __restrict__ out) {
  constexpr int V = 16 / (int)sizeof(T);                     // elements per 16-byte load: 4 floats / 2 doubles / 8 bf16
  __shared__ __align__(16) float tile_raw[kTrip + 4];
  const int tid = threadIdx.x;
  const int64_t a = blockIdx.x;
  const T* src = in + a * (int64_t)kTrip;
  const int head = (int)(((16u - (unsigned)((uintptr_t)src & 15u)) & 15u) / sizeof(T));   // elements before the first aligned 16 B
  float* tile = tile_raw + ((4 - (head & 3)) & 3);           // tile[head] is 16-byte aligned in shared memory
  if (tid < head) tile[tid] = ldf(src + tid);
  const int nvec = (kTrip - head) / V;
  if constexpr (sizeof(T) == 4) {
    const float4* v = reinterpret_cast<const float4*>(src + head);
    float4* t4 = reinterpret_cast<float4*>(tile + head);
#pragma unroll 4
    for (int i = tid; i < nvec; i += kPreThreads) t4[i] = __ldg(v + i);
  } else if constexpr (sizeof(T) == 2) {
    // bf16 rows packed on the host (btsb_host_pack_bf16): the alert starts at any 2-byte phase of a 16-byte line
    const uint4* v = reinterpret_cast<const uint4*>(src + head);
    float4* t4 = reinterpret_cast<float4*>(tile + head);
#pragma unroll 4
    for (int i = tid; i < nvec; i += kPreThreads) {
      const uint4 q = __ldg(v + i);
      t4[2 * i] = make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u),
                              __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
      t4[2 * i + 1] = make_float4(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xffff0000u),
                                  __uint_as_float(q.w << 16), __uint_as_float(q.w & 0xffff0000u));
    }
  } else {
    const double2* v = reinterpret_cast<const double2*>(src + head);
    float2* t2 = reinterpret_cast<float2*>(tile + head);
#pragma unroll 4
    for (int i = tid; i < nvec; i += kPreThreads) {
      const double2 d = __ldg(v + i);
      t2[i] = make_float2((float)d.x, (float)d.y);
    }
  }
  for (int i = head + nvec * V + tid; i < kTrip; i += kPreThreads) tile[i] = ldf(src + i);
  __syncthreads();
  float* dst = out + a * (int64_t)kTrip;
#pragma unroll 4
  for (int i = tid; i < kTrip; i += kPreThreads) {
    const int c = i / kPix, p = i - c * kPix;                // constants: multiply + shift
    dst[i] = tile[p * 3 + c];                                // stride-3 words across a warp: conflict-free
  }
}

// ---- make_triplet tail -----------------------------------------------------------------------------
__device__ __forceinline__ float nan_to_num_f32(float v) {
  if (isnan(v)) return 0.0f;
  if (isinf(v)) return v > 0 ? FLT_MAX : -FLT_MAX;
  return v;
}

template <typename TO>
__global__ void __launch_bounds__(kPreThreads)
pad_norm_kernel(const float* __restrict__ stamps, const int32_t* __restrict__ hw, int normalize,
                TO* __restrict__ out, uint8_t* __restrict__ drop_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);          // [63*63*3] HWC
  __shared__ double redd[kPreThreads / 32];
  __shared__ int redi[4][kPreThreads / 32];
  __shared__ int s_drop;
  __shared__ float s_nrm;
  __shared__ int s_norm_on;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t a = blockIdx.x;
  const float pad_val = 1e-9f;
  if (tid == 0) s_drop = 0;
  __syncthreads();

  for (int c = 0; c < 3; ++c) {
    const int h = min(max(hw[(a * 3 + c) * 2 + 0], 0), kImg), w = min(max(hw[(a * 3 + c) * 2 + 1], 0), kImg);
    const int cnt = h * w;
    const float* src = stamps + (a * 3 + c) * (int64_t)kPix;

    // pass 1: nanmedian == +-inf  <=>  position of the middle element(s) among the sorted non-NaN values
    int n_ok = 0, n_pinf = 0, n_ninf = 0, any_nz = 0;
    double sq = 0.0;
    for (int i = tid; i < cnt; i += kPreThreads) {
      const float v = __ldg(src + i);
      if (!isnan(v)) {
        ++n_ok;
        if (isinf(v)) { if (v > 0) ++n_pinf; else ++n_ninf; }
      }
      const float u = nan_to_num_f32(v);
      sq += (double)u * (double)u;
      any_nz |= (u != 0.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_ok += __shfl_xor_sync(0xffffffffu, n_ok, o);
      n_pinf += __shfl_xor_sync(0xffffffffu, n_pinf, o);
      n_ninf += __shfl_xor_sync(0xffffffffu, n_ninf, o);
      any_nz |= __shfl_xor_sync(0xffffffffu, any_nz, o);
    }
    sq = warp_sum(sq);
    if (lane == 0) { redi[0][wid] = n_ok; redi[1][wid] = n_pinf; redi[2][wid] = n_ninf; redi[3][wid] = any_nz; redd[wid] = sq; }
    __syncthreads();
    if (tid == 0) {
      int tn = 0, tp = 0, tm = 0, nz = 0; double ts = 0.0;
      for (int k = 0; k < kPreThreads / 32; ++k) { tn += redi[0][k]; tp += redi[1][k]; tm += redi[2][k]; nz |= redi[3][k]; ts += redd[k]; }
      bool med_inf = false;
      if (tn > 0) {
        if (tn & 1) {
          const int k = (tn - 1) / 2;
          med_inf = (k >= tn - tp) || (k < tm);
        } else {
          const int k1 = tn / 2 - 1, k2 = tn / 2;
          const bool hi_p = (k2 >= tn - tp), lo_m = (k1 < tm);
          // (+inf + -inf)/2 is NaN, which compares unequal to +-inf in the reference
          med_inf = (hi_p != lo_m);
        }
      }
      if (med_inf) s_drop = 1;
      const bool norm_on = normalize && !s_drop;          // `if normalize and not drop` (alert_utils.py:163)
      float nrm = 1.0f;
      if (norm_on) nrm = sqrtf((float)ts);                // float32 dot -> may overflow to +inf like numpy
      s_nrm = nrm;
      s_norm_on = norm_on ? 1 : 0;
      // all-zero test on the (possibly normalised) cutout: x/nrm == 0 for every pixel
      bool all_zero;
      if (norm_on) all_zero = (cnt > 0) && isinf(nrm);    // 0/0 = NaN != 0, so a zero stamp is NOT flagged here
      else all_zero = !nz;
      if (all_zero) s_drop = 1;
    }
    __syncthreads();
    const float nrm = s_nrm;
    const bool do_norm = s_norm_on != 0;
    // pass 2: normalise + pad into the staged HWC triplet
    for (int p = tid; p < kPix; p += kPreThreads) {
      const int y = p / kImg, x = p - y * kImg;
      float o = pad_val;
      if (y < h && x < w) {
        o = nan_to_num_f32(__ldg(src + y * w + x));
        if (do_norm) o = o / nrm;
      }
      tile[p * 3 + c] = o;
    }
    __syncthreads();
  }
  TO* dst = out + a * (int64_t)kTrip;
  for (int i = tid; i < kTrip; i += kPreThreads) dst[i] = (TO)tile[i];
  if (tid == 0) drop_out[a] = (uint8_t)s_drop;
}

// ---- training-time augmentation fused with the batch gather (utils.py:44-48, train.py:178-199) ----------------------
// out[b] = rot90^k( vflip?( hflip?( images[idx[b]] ))) -- pure index permutations, bit-exact.
// flags[b]: bit0 = horizontal flip, bit1 = vertical flip, bits 2-3 = k (counter-clockwise quarter turns).
__global__ void __launch_bounds__(256)
augment_gather_kernel(const float* __restrict__ images, const int64_t* __restrict__ idx, const uint8_t* __restrict__ flags,
                      float* __restrict__ out, int S) {
  const int64_t b = blockIdx.x;
  const int64_t src_img = idx[b];
  const int f = flags ? flags[b] : 0;
  const int hf = f & 1, vf = (f >> 1) & 1, k = (f >> 2) & 3;
  const int plane = S * S;
  const float* src = images + src_img * 3 * plane;
  float* dst = out + b * 3 * plane;
  for (int e = threadIdx.x; e < 3 * plane; e += 256) {
    const int c = e / plane, r = e - c * plane;
    const int i = r / S, j = r - i * S;
    int a, bb;
    if (k == 0) { a = i; bb = j; } else if (k == 1) { a = j; bb = S - 1 - i; }
    else if (k == 2) { a = S - 1 - i; bb = S - 1 - j; } else { a = S - 1 - j; bb = i; }
    if (vf) a = S - 1 - a;
    if (hf) bb = S - 1 - bb;
    dst[e] = __ldg(src + c * plane + a * S + bb);
  }
}

}  // namespace btsb

using namespace btsb;

extern "C" int btsb_augment_gather_f32(const float* images, const int64_t* idx, const uint8_t* flags, int64_t B, int S,
                                       float* out, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0 && S >= 1, "augment_gather: bad shape");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(images && idx && out, "augment_gather: null pointer");
  augment_gather_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(images, idx, flags, out, S);
  return launch_done("augment_gather");
}

extern "C" int btsb_preprocess_crop_norm(const void* in, int in_dtype, int64_t n, int crop_to_size, int normalize,
                                         int out_hwc, float* out, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(n >= 0, "crop_norm: n < 0");
  BTSB_REQUIRE(crop_to_size >= 1 && crop_to_size <= kImg, "crop_norm: crop_to_size %d not in [1,63]", crop_to_size);
  BTSB_REQUIRE(in_dtype == BTSB_F32 || in_dtype == BTSB_F64 || in_dtype == BTSB_BF16, "crop_norm: in_dtype must be F32, F64 or BF16");
  if (n == 0) return BTSB_OK;
  BTSB_REQUIRE(in && out, "crop_norm: null pointer");
  const int margin = (kImg - crop_to_size) / 2;
  cudaStream_t st = (cudaStream_t)stream;
  // bf16 input = rows packed by btsb_host_pack_bf16 for the scoring path: cast + transpose only
  BTSB_REQUIRE(in_dtype != BTSB_BF16 || (crop_to_size == kImg && !normalize && !out_hwc),
               "crop_norm: BF16 input is the cast + transpose path only (crop 63, no normalisation, NCHW)");
  if (crop_to_size == kImg && !normalize && !out_hwc) {
    if (in_dtype == BTSB_BF16)
      cast_transpose63_kernel<__nv_bfloat16><<<(unsigned)n, kPreThreads, 0, st>>>((const __nv_bfloat16*)in, out);
    else if (in_dtype == BTSB_F32) cast_transpose63_kernel<float><<<(unsigned)n, kPreThreads, 0, st>>>((const float*)in, out);
    else cast_transpose63_kernel<double><<<(unsigned)n, kPreThreads, 0, st>>>((const double*)in, out);
    return launch_done("crop_norm");
  }
  if (in_dtype == BTSB_F32) {
    const int smem = kTrip * 4;
    crop_norm_kernel<float><<<(unsigned)n, kPreThreads, smem, st>>>((const float*)in, crop_to_size, margin, normalize, out_hwc, out);
  } else {
    const int smem = kTrip * 8;
    static std::atomic<uint64_t> attr{0};
    if (first_use_on_device(attr)) {
      BTSB_CUDA(cudaFuncSetAttribute(crop_norm_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "crop_norm attr");
    }
    crop_norm_kernel<double><<<(unsigned)n, kPreThreads, smem, st>>>((const double*)in, crop_to_size, margin, normalize, out_hwc, out);
  }
  return launch_done("crop_norm");
}

extern "C" int btsb_preprocess_pad_norm(const float* stamps, const int32_t* hw, int64_t n, int normalize, void* out,
                                        int out_dtype, uint8_t* drop, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(n >= 0, "pad_norm: n < 0");
  BTSB_REQUIRE(out_dtype == BTSB_F32 || out_dtype == BTSB_F64, "pad_norm: out_dtype must be F32 or F64");
  if (n == 0) return BTSB_OK;
  BTSB_REQUIRE(stamps && hw && out && drop, "pad_norm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = kTrip * 4;
  if (out_dtype == BTSB_F64)
    pad_norm_kernel<double><<<(unsigned)n, kPreThreads, smem, st>>>(stamps, hw, normalize, (double*)out, drop);
  else
    pad_norm_kernel<float><<<(unsigned)n, kPreThreads, smem, st>>>(stamps, hw, normalize, (float*)out, drop);
  return launch_done("pad_norm");
}
