// conv0 of the CPC gEncoder + ChannelNorm + ReLU on the tensor cores (tf32), bf16 output.
// Reference: vap/encoder_components.py:83-84,99 (Conv1d(1,256,k=10,s=5,p=3)), :62-70
// (ChannelNorm: unbiased variance over the 256 channels of one time step).
//
// The layer writes 512 bytes per frame (65.5 MB per 20 s chunk in bf16) for 2560 MACs, so it
// has to run at HBM write speed; on CUDA cores it is FMA-issue bound at half of that
// (the round-1 FFMA2 kernel, removed). Here the conv is a GEMM with K = 16:
//   A[128 frames][16]  = im2col of the waveform (samples 5f-3 .. 5f+6, then 1.0, then zeros), tf32,
//                        built in shared memory by 4 warps (the row stride of 5 samples is not
//                        expressible as a TMA stride)
//   B[256 ch][16]      = g_c (w_c - wbar) | g_c (b_c - bbar) | 0 : the ChannelNorm mean and affine scale
//                        are folded into the weights on the host (conv0_v2_fold), resident in smem
//   D[128][256] fp32   in TMEM (two accumulators)
// and the per-frame 1/sqrt(var + eps) is the closed form x'Gx + 2h.x + s evaluated by the thread
// that builds the frame's A row. The epilogue is one FMA per output (acc * rstd_f + beta_c), ReLU,
// bf16 pack into a swizzled staging tile and a TMA store. Every epilogue warp stages and stores its own 32 rows x 64
// channels (4 KB, two buffers): no barrier couples the warps, so their TMEM loads, math and stores overlap freely
// (with one 128-row store per column block and two named barriers per block the tile took ~3500 clk, latency-bound).
#include <string>

#include "common.cuh"
#include "tc_common.cuh"

namespace vapb {

using namespace tc;

namespace {

constexpr int CT_SPLIT = 4;                      // column groups of the epilogue: 4 warps (TMEM lane quadrants) each
constexpr int CT_EPI_WARPS = 4 * CT_SPLIT;
constexpr int CT_CHUNKS = 8 / CT_SPLIT;          // 32-column chunks per epilogue thread
constexpr int CT_THREADS = (5 + CT_EPI_WARPS) * 32;  // warps 0-3 build A, warp 4 issues MMAs (and owns TMEM), then epilogue
constexpr int CT_A_BYTES = 4 * 128 * 16;   // [k chunk of 4][128 rows][16 B]
constexpr int CT_W_BYTES = 4 * 256 * 16;   // [k chunk of 4][256 ch][16 B]
constexpr int CT_XS = 648;                 // samples of one tile (645 used)
constexpr int CT_OFF_A = 0;                         // [2]
constexpr int CT_OFF_W = 2 * CT_A_BYTES;
static_assert(CT_CHUNKS == 2, "the staging tile is 32 rows x 128 B (64 channels per warp), SWIZZLE_128B");
constexpr int CT_OFF_STG = CT_OFF_W + CT_W_BYTES;   // [epilogue warp][2] x (32 rows x 128 B)
constexpr int CT_OFF_XS = CT_OFF_STG + CT_EPI_WARPS * 8192;    // float [2][CT_XS]
constexpr int CT_OFF_RS = CT_OFF_XS + 2 * CT_XS * 4;  // float [4][128]
constexpr int CT_OFF_BE = CT_OFF_RS + 4 * 128 * 4;    // float [256]
constexpr int CT_OFF_BAR = CT_OFF_BE + 256 * 4;
constexpr int CT_SMEM = CT_OFF_BAR + 128 + 1024 /*alignment slack*/;

struct alignas(64) Conv0TcParams {
  CUtensorMap tma_out;  // (256, L0, nseq) bf16 over the padded activation buffer, box (64, 32, 1), SW128
  const float* wav;
  const float* u;     // [10][256] folded taps
  const float* d;     // [256] folded offsets
  const float* beta;  // [256]
  Conv0Stats cs;
  int batch, seq0, nseq, tiles_per_seq;
  long long n_samples, L0;
  int fp16;
};

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {  // round to nearest (the MMA itself would truncate)
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}
// kind::tf32 instruction descriptor: a/b format 2 = TF32, fp32 accumulate, K-major operands
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(CT_THREADS, 1) conv0_tc_kernel(const __grid_constant__ Conv0TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + CT_OFF_BAR;
  auto a_full = [&](int b) { return bar_base + 8u * b; };
  auto a_empty = [&](int b) { return bar_base + 8u * (2 + b); };
  auto tfull = [&](int a) { return bar_base + 8u * (4 + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (6 + a); };
  const uint32_t tmem_slot = bar_base + 8u * 8;
  float* xs = reinterpret_cast<float*>(smem_gen + CT_OFF_XS);
  float* rs = reinterpret_cast<float*>(smem_gen + CT_OFF_RS);
  float* be = reinterpret_cast<float*>(smem_gen + CT_OFF_BE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tma_out);
    for (int b = 0; b < 2; ++b) {
      mbar_init(a_full(b), 128);
      mbar_init(a_empty(b), 1);
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), CT_EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  // resident operands: W (tf32, unswizzled K-major core matrices), beta, and the constant parts of both A tiles
  for (int i = threadIdx.x; i < 256 * 16; i += CT_THREADS) {
    const int n = i >> 4, k = i & 15;
    const float v = k < 10 ? p.u[k * kDim + n] : (k == 10 ? p.d[n] : 0.f);
    *reinterpret_cast<float*>(smem_gen + CT_OFF_W + (k >> 2) * 4096 + n * 16 + (k & 3) * 4) = to_tf32(v);
  }
  for (int i = threadIdx.x; i < 256; i += CT_THREADS) be[i] = p.beta[i];
  for (int i = threadIdx.x; i < 2 * 128; i += CT_THREADS)  // k chunk 3 (taps 12..15) of both A tiles is zero
    *reinterpret_cast<float4*>(smem_gen + CT_OFF_A + (i >> 7) * CT_A_BYTES + 3 * 2048 + (i & 127) * 16) =
        make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // above: weights only

  const int num_tiles = p.nseq * p.tiles_per_seq;
  auto wav_of = [&](int tile, long long* s0) -> const float* {
    const int lseq = tile / p.tiles_per_seq, fb = tile % p.tiles_per_seq;
    const int seq = p.seq0 + lseq;  // channel-major sequence id: c * batch + item
    const int ch = seq / p.batch, item = seq % p.batch;
    *s0 = 5LL * 128 * fb - 3;
    return p.wav + ((long long)item * 2 + ch) * p.n_samples;
  };

  if (warp < 4) {
    // ===== builders: thread r owns frame r of the tile
    const int r = threadIdx.x;
    auto fetch = [&](int tile, float (&reg)[6]) {
      long long s0;
      const float* x = wav_of(tile, &s0);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int i = r + j * 128;
        const long long s = s0 + i;
        reg[j] = (i < 645 && s >= 0 && s < p.n_samples) ? __ldg(x + s) : 0.f;
      }
    };
    auto stash = [&](int buf, const float (&reg)[6]) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int i = r + j * 128;
        if (i < CT_XS) xs[buf * CT_XS + i] = reg[j];
      }
    };
    auto bar_build = [&]() { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    float reg[6];
    int it = 0;
    if ((int)blockIdx.x < num_tiles) {
      fetch(blockIdx.x, reg);
      stash(0, reg);
    }
    bar_build();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int nxt = tile + gridDim.x;
      if (nxt < num_tiles) fetch(nxt, reg);  // in flight while this tile is built
      float xv[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) xv[k] = xs[buf * CT_XS + 5 * r + k];
      // 1 / sqrt(var + eps) of the frame from the closed form
      float ss = p.cs.s;
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        float y = p.cs.h2[k];
#pragma unroll
        for (int l = 0; l < 10; ++l) y = fmaf(p.cs.G[k][l], xv[l], y);
        ss = fmaf(xv[k], y, ss);
      }
      const float rstd = rsqrtf(fmaxf(ss, 0.f) * (1.0f / (kDim - 1)) + kEps);
      mbar_wait(a_empty(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      // rs is a ring of 4 tiles. Stored only after a_empty(it): MMA(it-2) has completed, so the epilogue released the
      // accumulator of tile it-4 and is done with this slot. (Stored before the wait, the slot could still be unread by
      // a late epilogue warp of tile it-4: a rare, timing-dependent wrong rstd for 32 frames.)
      rs[(it & 3) * 128 + r] = rstd;
      uint8_t* arow = smem_gen + CT_OFF_A + buf * CT_A_BYTES + r * 16;
      *reinterpret_cast<float4*>(arow) = make_float4(to_tf32(xv[0]), to_tf32(xv[1]), to_tf32(xv[2]), to_tf32(xv[3]));
      *reinterpret_cast<float4*>(arow + 2048) = make_float4(to_tf32(xv[4]), to_tf32(xv[5]), to_tf32(xv[6]), to_tf32(xv[7]));
      *reinterpret_cast<float4*>(arow + 4096) = make_float4(to_tf32(xv[8]), to_tf32(xv[9]), 1.0f, 0.f);
      fence_proxy_async();
      mbar_arrive(a_full(buf));
      if (nxt < num_tiles) stash(buf ^ 1, reg);
      bar_build();
    }
  } else if (warp == 4) {
    // ===== MMA issuer: two K = 8 tf32 steps per tile
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, 256);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty(buf), ph ^ 1u);
        mbar_wait(a_full(buf), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + CT_OFF_A + buf * CT_A_BYTES, w_addr = smem_base + CT_OFF_W;
#pragma unroll
        for (int s = 0; s < 2; ++s)
          umma_tf32(tmem_base + buf * 256, make_smem_desc_nosw(a_addr + s * 4096, 2048, 128),
                    make_smem_desc_nosw(w_addr + s * 8192, 4096, 128), idesc, s != 0);
        umma_commit(a_empty(buf));
        umma_commit(tfull(buf));
      }
    }
  } else {
    // ===== epilogue: thread = (accumulator row, column group); each warp stores its own 32 rows
    const int quad = warp & 3, grp = (warp - 5) >> 2;
    const int row = quad * 32 + lane;
    const int cbase = grp * (32 * CT_CHUNKS);
    const uint32_t stg_addr = smem_base + CT_OFF_STG + (uint32_t)(warp - 5) * 8192u;
    uint8_t* stg_gen = smem_gen + CT_OFF_STG + (warp - 5) * 8192;
    const uint32_t sw128 = (uint32_t)(lane & 7);
    uint32_t stg_cnt = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int lseq = tile / p.tiles_per_seq, f0 = (tile % p.tiles_per_seq) * 128;
      mbar_wait(tfull(buf), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const float rstd = rs[(it & 3) * 128 + row];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256 + cbase;
      uint32_t r[2][32];
      tmem_ld32(taddr, r[0]);
      // warp staging tile: 32 rows x 128 B (this warp's 64 channels), SWIZZLE_128B (16-byte chunk j of row r at
      // r*128 + ((j ^ (r & 7)) << 4)), two buffers: the store that read this buffer two tiles ago must have drained it
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
      const uint32_t boff = (stg_cnt & 1u) * 4096u;
      uint8_t* rowp = stg_gen + boff + (uint32_t)lane * 128u;
#pragma unroll
      for (int c = 0; c < CT_CHUNKS; ++c) {
        tmem_ld_wait();
        if (c + 1 < CT_CHUNKS) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&be[cbase + c * 32 + i]);
          const float v0 = fmaf(__uint_as_float(r[c & 1][i]), rstd, b.x);
          const float v1 = fmaf(__uint_as_float(r[c & 1][i + 1]), rstd, b.y);
          const float v2 = fmaf(__uint_as_float(r[c & 1][i + 2]), rstd, b.z);
          const float v3 = fmaf(__uint_as_float(r[c & 1][i + 3]), rstd, b.w);
          pk[i >> 1] = pack16(fmaxf(v0, 0.f), fmaxf(v1, 0.f), p.fp16);
          pk[(i >> 1) + 1] = pack16(fmaxf(v2, 0.f), fmaxf(v3, 0.f), p.fp16);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(rowp + (((uint32_t)(c * 4 + j) ^ sw128) << 4)) =
              make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&p.tma_out, stg_addr + boff, cbase, f0 + quad * 32, lseq);
        bulk_commit();
      }
      ++stg_cnt;
      tc_fence_before();
      mbar_arrive(tempty(buf));
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int launch_conv0_tc(cudaStream_t st, const float* wav, int batch, long long n_samples, int seq0, int nseq,
                    long long L0, const float* u, const float* d, const float* beta, const Conv0Stats& cs,
                    __nv_bfloat16* out, long long out_seq_stride, int out_pad_rows, int n_sm, std::string* err) {
  Conv0TcParams p{};
  {
    const uint64_t dims[3] = {(uint64_t)kDim, (uint64_t)L0, (uint64_t)nseq};
    const uint64_t strides[2] = {(uint64_t)kDim, (uint64_t)out_seq_stride};
    const uint32_t box[3] = {64, 32, 1};
    if (!make_tmap(&p.tma_out, out + (long long)out_pad_rows * kDim, 2, 3, dims, strides, box, 128, err)) return -1;
  }
  p.wav = wav;
  p.u = u;
  p.d = d;
  p.beta = beta;
  p.cs = cs;
  p.batch = batch;
  p.seq0 = seq0;
  p.nseq = nseq;
  p.tiles_per_seq = (int)((L0 + 127) / 128);
  p.n_samples = n_samples;
  p.L0 = L0;
  p.fp16 = g_fp16;
  static bool configured_on[64] = {};  // the attribute is per device (one process may drive several GPUs)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_on[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM) != cudaSuccess) {
      if (err) *err = "conv0_tc: cannot reserve shared memory";
      return -1;
    }
    configured = true;
  }
  const long long tiles = (long long)nseq * p.tiles_per_seq;
  const int grid = tiles < n_sm ? (int)tiles : n_sm;
  launch_pdl(conv0_tc_kernel, grid, CT_THREADS, CT_SMEM, st, p);
  return 1;
}

}  // namespace vapb
