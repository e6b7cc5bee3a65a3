// Nearest-code search on the 5th-gen tensor cores (sm_100a).
//
// Replaces the reference's   dist = -cdist(x, embed) | einsum(x, embed) ; ind = dist.argmax(-1)
// (vector_quantize_pytorch.py:58-62, :741-747, :130-145) without materialising the (N x K) matrix.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer : x tile (A, 128 rows, stationary in smem for the whole code sweep) and
//                              codebook tiles (B, BN codes x 64 dims per stage) -> 128B-swizzled smem
//   warp 1      MMA issuer   : tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32) into TMEM;
//                              split-precision passes (a0,c_hi)+(a0,c_lo)[+(a1,c_hi)] accumulate into
//                              the SAME accumulator, two accumulator stages (2 x 256 TMEM columns)
//   warps 2..5  epilogue     : tcgen05.ld (lane == row, so a thread owns a whole row of scores),
//                              score = acc - bias, running arg-max with an error-band candidate list
//
// Exactness: the tensor-core score of a (row, code) pair differs from the exact fp32 value by at most
// tau = margin_rel * ||x|| * max||c||.  A row is certified when its best score leads every other score
// by more than W = 2*tau; otherwise (row, candidates) goes to `flagged` and vqb_fix_flagged re-scores
// it with the reference's exact formula.  The band test is conservative (may over-flag, never under-flag).
#include "ptx.cuh"
#include "vqb_common.cuh"

namespace vqb {

constexpr int BM = 128;         // rows of x per tile (UMMA M, one TMEM lane per row)
constexpr int BK = 64;          // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;      // K of one tcgen05.mma for 16-bit inputs
constexpr int A_SUB_BYTES = BM * BK * 2;  // 16 KiB: one (plane, k-block) sub-tile of A
constexpr int MAX_A_SUB = 8;    // n_a * ceil(D/64) <= 8  -> A <= 128 KiB
constexpr int MAX_STAGES = 6;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_CTRL_BYTES = 1024;  // barriers + tmem ptr + row norms
constexpr int SMEM_LIMIT = 232448;     // 227 KiB opt-in maximum per CTA

struct AssignParams {
  int64_t N;
  int D, K, Kpad, BN;
  int n_a, n_passes;   // passes: 0:(a0,hi) 1:(a0,lo) 2:(a1,hi)
  int KB;              // ceil(D / 64)
  int n_stages;
  int num_row_tiles, num_code_tiles;
  float margin_rel;
  const float* bias;   // [Kpad]
  const float* cmax;   // [1]
  int32_t* idx;
  vqb_flag_entry* flagged;
  int32_t* flag_count;
  float* dbg_best;
};

struct Ctrl {  // lives at the start of dynamic smem
  uint64_t a_full, a_empty;
  uint64_t b_full[MAX_STAGES], b_empty[MAX_STAGES];
  uint64_t t_full[2], t_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float xn2[BM];
};
static_assert(sizeof(Ctrl) <= SMEM_CTRL_BYTES, "control block too large");

// Candidate-band update for one score (slow path; entered for ~ln(K) elements per row).
struct RowState {
  float best, thr, W;
  int i0, i1, n;
  __device__ __forceinline__ void update(float v, int c) {
    if (v > thr) {
      if (v > best) {
        if (v - best > W) { n = 1; } else { n += 1; i1 = i0; }
        best = v;
        i0 = c;
      } else {
        n += 1;
        i1 = c;
      }
      thr = best - W;
    }
  }
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
vq_assign_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const AssignParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = (smem_base + SMEM_CTRL_BYTES + 1023u) & ~1023u;    // swizzle-128B tiles need 1024 B alignment
  const uint8_t* a_gen = smem + (a_base - smem_base);                         // same place, generic address
  const uint32_t b_base = a_base + p.n_a * p.KB * A_SUB_BYTES;               // 1024-aligned
  const uint32_t b_stage_bytes = p.BN * BK * 2;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ------------------------------------------------------------------ one-time setup
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    mbar_init(smem_u32(&ctrl->a_full), 1);
    mbar_init(smem_u32(&ctrl->a_empty), 1 + 4);  // MMA commit + 4 epilogue warps (they read A for the row norms)
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(smem_u32(&ctrl->b_full[s]), 1);
      mbar_init(smem_u32(&ctrl->b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&ctrl->t_full[s]), 1);
      mbar_init(smem_u32(&ctrl->t_empty[s]), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  const int my_tiles = (p.num_row_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = blockIdx.x + t * gridDim.x;
        const int row0 = tile * BM;
        mbar_wait(smem_u32(&ctrl->a_empty), (t & 1) ^ 1);
        mbar_arrive_expect_tx(smem_u32(&ctrl->a_full), p.n_a * p.KB * A_SUB_BYTES);
        for (int pl = 0; pl < p.n_a; ++pl)
          for (int kb = 0; kb < p.KB; ++kb)
            tma_load_3d(a_base + (pl * p.KB + kb) * A_SUB_BYTES, &tmA, smem_u32(&ctrl->a_full), kb * BK, row0, pl);
        for (int ct = 0; ct < p.num_code_tiles; ++ct) {
          for (int ps = 0; ps < p.n_passes; ++ps) {
            const int bplane = (ps == 1) ? 1 : 0;
            for (int kb = 0; kb < p.KB; ++kb) {
              mbar_wait(smem_u32(&ctrl->b_empty[stage]), ph ^ 1);
              mbar_arrive_expect_tx(smem_u32(&ctrl->b_full[stage]), b_stage_bytes);
              tma_load_3d(b_base + stage * b_stage_bytes, &tmB, smem_u32(&ctrl->b_full[stage]), kb * BK, ct * p.BN, bplane);
              if (++stage == p.n_stages) { stage = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, p.BN);
      int stage = 0;
      uint32_t ph = 0;
      uint32_t it = 0;  // accumulator iteration counter (across row tiles)
      for (int t = 0; t < my_tiles; ++t) {
        mbar_wait(smem_u32(&ctrl->a_full), t & 1);
        tc_fence_after();
        for (int ct = 0; ct < p.num_code_tiles; ++ct, ++it) {
          const uint32_t as = it & 1;
          mbar_wait(smem_u32(&ctrl->t_empty[as]), ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * 256;
          uint32_t acc = 0;
          for (int ps = 0; ps < p.n_passes; ++ps) {
            const int aplane = (ps == 2) ? 1 : 0;
            for (int kb = 0; kb < p.KB; ++kb) {
              mbar_wait(smem_u32(&ctrl->b_full[stage]), ph);
              tc_fence_after();
              const uint32_t a_addr = a_base + (aplane * p.KB + kb) * A_SUB_BYTES;
              const uint32_t b_addr = b_base + stage * b_stage_bytes;
              const int rem = p.D - kb * BK;
              const int ksteps = rem >= BK ? (BK / UMMA_K) : (rem + UMMA_K - 1) / UMMA_K;
              for (int k = 0; k < ksteps; ++k) {
                umma_bf16_ss(d_tmem, umma_smem_desc_sw128(a_addr + k * UMMA_K * 2),
                             umma_smem_desc_sw128(b_addr + k * UMMA_K * 2), idesc, acc);
                acc = 1;
              }
              umma_commit(smem_u32(&ctrl->b_empty[stage]));  // stage reusable once these MMAs retire
              if (++stage == p.n_stages) { stage = 0; ph ^= 1; }
            }
          }
          umma_commit(smem_u32(&ctrl->t_full[as]));  // accumulator complete -> epilogue
        }
        umma_commit(smem_u32(&ctrl->a_empty));  // all MMAs reading this A tile retired
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5)
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int row_in_tile = lg * 32 + lane;  // TMEM lane == row of the tile
    const float cmax = __ldg(p.cmax);
    uint32_t it = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      // ---- row norms from the A tile in smem (conflict-free: a warp reads 4 full 128 B rows per request)
      mbar_wait(smem_u32(&ctrl->a_full), t & 1);
      {
        const int sub = lane >> 3, chunk = lane & 7;
        for (int i = 0; i < 8; ++i) {
          const int r = lg * 32 + i * 4 + sub;
          const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
          float acc2 = 0.f;
          for (int kb = 0; kb < p.KB; ++kb) {
            float v[8];
            {
              uint4 u = *reinterpret_cast<const uint4*>(a_gen + kb * A_SUB_BYTES + off);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[2 * e] = __uint_as_float(w[e] << 16);
                v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
              }
            }
            if (p.n_a == 2) {
              uint4 u = *reinterpret_cast<const uint4*>(a_gen + (p.KB + kb) * A_SUB_BYTES + off);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[2 * e] += __uint_as_float(w[e] << 16);
                v[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) acc2 = fmaf(v[e], v[e], acc2);
          }
          acc2 += __shfl_xor_sync(0xffffffffu, acc2, 1);
          acc2 += __shfl_xor_sync(0xffffffffu, acc2, 2);
          acc2 += __shfl_xor_sync(0xffffffffu, acc2, 4);
          if (chunk == 0) ctrl->xn2[r] = acc2;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctrl->a_empty));
      }
      RowState st;
      st.W = 2.f * p.margin_rel * sqrtf(ctrl->xn2[row_in_tile]) * cmax + 1e-30f;
      st.best = -INFINITY;
      st.thr = -INFINITY;
      st.i0 = 0;
      st.i1 = -1;
      st.n = 0;
      __syncwarp();  // xn2 reads done before the next tile's writers (same warp) run

      for (int ct = 0; ct < p.num_code_tiles; ++ct, ++it) {
        const uint32_t as = it & 1;
        mbar_wait(smem_u32(&ctrl->t_full[as]), (it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + as * 256;
        const int code0 = ct * p.BN;
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          if (p.BN - c0 >= 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_addr + c0, r);
            tmem_wait_ld();
            const float4* bp = reinterpret_cast<const float4*>(p.bias + code0 + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = __ldg(bp + j);
              const float v0 = __uint_as_float(r[4 * j + 0]) - bb.x;
              const float v1 = __uint_as_float(r[4 * j + 1]) - bb.y;
              const float v2 = __uint_as_float(r[4 * j + 2]) - bb.z;
              const float v3 = __uint_as_float(r[4 * j + 3]) - bb.w;
              const float m = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
              if (m > st.thr) {
                const int c = code0 + c0 + 4 * j;
                st.update(v0, c);
                st.update(v1, c + 1);
                st.update(v2, c + 2);
                st.update(v3, c + 3);
              }
            }
          } else {  // BN is a multiple of 16: one 16-column tail
            uint32_t r[16];
            tmem_ld_32x32b_x16(t_addr + c0, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) st.update(__uint_as_float(r[j]) - __ldg(p.bias + code0 + c0 + j), code0 + c0 + j);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctrl->t_empty[as]));
      }
      const int64_t row = static_cast<int64_t>(tile) * BM + row_in_tile;
      if (row < p.N) {
        p.idx[row] = st.i0;
        if (p.dbg_best) p.dbg_best[row] = st.best;
        if (st.n >= 2) {
          const int slot = atomicAdd(p.flag_count, 1);
          vqb_flag_entry e;
          e.row = static_cast<int32_t>(row);
          e.cand0 = st.i0;
          e.cand1 = st.i1;
          e.count = st.n;
          p.flagged[slot] = e;
        }
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

// bf16 tensor [planes][rows][D] (row-major) -> boxes of {64 dims, box_rows rows, 1 plane}, 128B swizzle,
// out-of-bounds elements read as zero (ragged N / K / D are handled by the zero fill).
static int make_map(CUtensorMap* m, const void* base, int D, int64_t rows, int planes, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return VQB_E_DRIVER;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(D) * 2, static_cast<cuuint64_t>(D) * 2 * static_cast<cuuint64_t>(rows)};
  cuuint32_t box[3] = {BK, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQB_OK : VQB_E_DRIVER;
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_padded_codes(int K) {
  if (K <= 0) return 0;
  const int BN = code_tile(K);
  return (K + BN - 1) / BN * BN;
}

extern "C" int vqb_assign(const void* a_planes, int n_a, int64_t N, int D, const void* b_planes, const float* bias,
                          const float* cmax, int K, float margin_rel, int n_passes, int32_t* idx,
                          vqb_flag_entry* flagged, int32_t* flag_count, float* dbg_best, void* stream) {
  if (!a_planes || !b_planes || !bias || !cmax || !idx || !flagged || !flag_count) return VQB_E_INVALID;
  if (N <= 0 || D <= 0 || K <= 0 || (n_a != 1 && n_a != 2)) return VQB_E_INVALID;
  if (n_passes == 0) n_passes = (n_a == 2) ? 3 : 2;
  if (n_passes < 1 || n_passes > 3 || (n_passes == 3 && n_a != 2)) return VQB_E_INVALID;
  if (D % 8 != 0) return VQB_E_UNSUPPORTED;
  const int KB = (D + BK - 1) / BK;
  if (n_a * KB > MAX_A_SUB) return VQB_E_UNSUPPORTED;
  if (N > (static_cast<int64_t>(1) << 31) - BM) return VQB_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a_planes) | reinterpret_cast<uintptr_t>(b_planes) | reinterpret_cast<uintptr_t>(bias)) & 15)
    return VQB_E_ALIGN;
  int rc = check_device();
  if (rc) return rc;

  AssignParams p;
  p.N = N; p.D = D; p.K = K;
  p.BN = code_tile(K);
  p.Kpad = vqb_padded_codes(K);
  p.n_a = n_a; p.n_passes = n_passes; p.KB = KB;
  p.num_row_tiles = static_cast<int>((N + BM - 1) / BM);
  p.num_code_tiles = p.Kpad / p.BN;
  p.margin_rel = margin_rel;
  p.bias = bias; p.cmax = cmax; p.idx = idx; p.flagged = flagged; p.flag_count = flag_count; p.dbg_best = dbg_best;
  const int a_bytes = n_a * KB * A_SUB_BYTES;
  const int b_stage = p.BN * BK * 2;
  int stages = (SMEM_LIMIT - SMEM_CTRL_BYTES - 1024 /*align slack*/ - a_bytes) / b_stage;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return VQB_E_UNSUPPORTED;
  p.n_stages = stages;
  const int smem_bytes = SMEM_CTRL_BYTES + 1024 + a_bytes + stages * b_stage;

  CUtensorMap tmA, tmB;
  rc = make_map(&tmA, a_planes, D, N, n_a, BM);
  if (rc) return rc;
  rc = make_map(&tmB, b_planes, D, p.Kpad, 2, p.BN);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vq_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  int grid = p.num_row_tiles < num_sms() ? p.num_row_tiles : num_sms();
  vq_assign_kernel<<<grid, NUM_THREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  return static_cast<int>(cudaGetLastError());
}
