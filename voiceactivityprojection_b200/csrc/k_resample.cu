// Rational polyphase resampler on the device: the input path in front of the model
// (vap/audio.py:65-68 calls torchaudio.functional.resample).
//
// torchaudio's algorithm (functional.py, _apply_sinc_resample_kernel): with orig/new reduced by their gcd, pad the row
// with `width` zeros on the left and `width + orig` on the right, then out[n*new + p] = sum_k bank[p][k] * x_pad[n*orig + k]
// for k < K = 2*width + orig, cut to ceil(new * n_in / orig) samples. The windowed-sinc bank is computed by the host
// (voiceactivityprojection_b200/audio.py) and passed in. HBM-bound byte work: one thread per output sample, bank in
// shared memory when it fits, input rows read through L1 (neighbouring outputs share all but `orig` of their taps).
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

namespace vapb {

namespace {

struct ResampleParams {
  const void* x;
  long long item_stride, chan_stride, elem_stride;
  int channels;
  long long n_in, n_out, out_row_stride;
  int orig, new_, width, K;
  const float* bank;
  float* out;
};

template <typename TIn>
__device__ __forceinline__ float load_sample(const TIn* p);
template <>
__device__ __forceinline__ float load_sample<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* p) {
  return (float)__ldg(p) * (1.0f / 32768.0f);  // torchaudio.load's int16 normalisation
}

template <typename TIn, bool SMEM_BANK>
__global__ void __launch_bounds__(256) resample_kernel(const ResampleParams p) {
  extern __shared__ float bank_s[];
  if (SMEM_BANK) {
    for (int i = threadIdx.x; i < p.new_ * p.K; i += blockDim.x) bank_s[i] = p.bank[i];
    __syncthreads();
  }
  const long long row = blockIdx.y;
  const long long item = row / p.channels, ch = row % p.channels;
  const TIn* x = reinterpret_cast<const TIn*>(p.x) + item * p.item_stride + ch * p.chan_stride;
  float* out = p.out + row * p.out_row_stride;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < p.n_out;
       m += (long long)gridDim.x * blockDim.x) {
    const long long n = m / p.new_;
    const int ph = (int)(m - n * p.new_);
    const float* h = (SMEM_BANK ? bank_s : p.bank) + ph * p.K;
    const long long i0 = n * p.orig - p.width;
    int k0 = i0 < 0 ? (int)-i0 : 0;
    int k1 = p.K;
    if (i0 + k1 > p.n_in) k1 = (int)(p.n_in - i0);
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(SMEM_BANK ? h[k] : __ldg(h + k), load_sample<TIn>(x + (i0 + k) * p.elem_stride), acc);
    out[m] = acc;
  }
}

}  // namespace

// x_fmt: 0 = float32, 1 = int16 PCM. Returns launches (1) or -1 with *err set.
int launch_resample(cudaStream_t st, const void* x, int x_fmt, long long items, int channels, long long n_in,
                    long long item_stride, long long chan_stride, long long elem_stride, int orig, int new_, int width,
                    const float* bank, float* out, long long n_out, long long out_row_stride, std::string* err) {
  if (items < 0 || channels < 1 || n_in < 0 || orig < 1 || new_ < 1 || width < 0 || (x_fmt != 0 && x_fmt != 1)) {
    if (err) *err = "resample: invalid argument";
    return -1;
  }
  const long long want = (new_ * n_in + orig - 1) / orig;
  if (n_out > want || n_out < 0) {
    if (err) *err = "resample: n_out exceeds ceil(new * n_in / orig)";
    return -1;
  }
  const long long rows = items * channels;
  if (rows == 0 || n_out == 0) return 0;
  if (rows > 65535) {
    if (err) *err = "resample: more than 65535 rows in one call";
    return -1;
  }
  ResampleParams p{};
  p.x = x; p.item_stride = item_stride; p.chan_stride = chan_stride; p.elem_stride = elem_stride;
  p.channels = channels; p.n_in = n_in; p.n_out = n_out; p.out_row_stride = out_row_stride;
  p.orig = orig; p.new_ = new_; p.width = width; p.K = 2 * width + orig;
  p.bank = bank; p.out = out;
  const size_t bank_bytes = (size_t)new_ * p.K * sizeof(float);
  const bool smem_bank = bank_bytes <= 40 * 1024;
  long long bx = (n_out + 255) / 256;
  if (bx > 4096) bx = 4096;
  const dim3 grid((unsigned)bx, (unsigned)rows);
  if (x_fmt == 0) {
    if (smem_bank) resample_kernel<float, true><<<grid, 256, bank_bytes, st>>>(p);
    else resample_kernel<float, false><<<grid, 256, 0, st>>>(p);
  } else {
    if (smem_bank) resample_kernel<int16_t, true><<<grid, 256, bank_bytes, st>>>(p);
    else resample_kernel<int16_t, false><<<grid, 256, 0, st>>>(p);
  }
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("resample launch: ") + cudaGetErrorString(e);
    return -1;
  }
  return 1;
}

}  // namespace vapb
