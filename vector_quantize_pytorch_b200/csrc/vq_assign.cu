// Nearest-code search on the 5th-gen tensor cores (sm_100a).
//
// Replaces the reference's   dist = -cdist(x, embed) | einsum(x, embed) ; ind = dist.argmax(-1)
// (vector_quantize_pytorch.py:58-62, :741-747, :130-145) without materialising the (N x K) matrix.
//
// One persistent CTA per SM, warp-specialised (10 warps):
//   warp 0      TMA producer : x tile (A: 128 rows, stationary in smem for the whole code sweep, refilled
//                              k-block by k-block as the last sweep of the previous tile releases it),
//                              codebook tiles (B: BN codes x 64 dims per stage) and the per-tile bias
//                              block (Bext: BN codes x 16) -> swizzled smem
//   warp 1      MMA issuer   : tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32) into TMEM.
//                              Per code tile: one K=16 MMA  [1 1 1 0..] x [-b1 -b2 -b3 0..]^T  that seeds
//                              the accumulator with -0.5||c||^2 (three bf16 terms = exact fp32), then the
//                              split-precision passes (a0,c_hi)+(a0,c_lo)[+(a1,c_hi)] accumulate on top.
//                              Two accumulator stages (2 x 256 TMEM columns).
//   warps 2..9  epilogue     : tcgen05.ld (lane == row).  Warps w and w+4 share a TMEM lane group and split
//                              the columns; a thread keeps the running arg-max of its row slice plus an
//                              error-band candidate list, the two slices are merged once per row tile.
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
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int AEXT_BYTES = BM * 32;       // [128 rows][16 bf16], 32-byte swizzle
constexpr int SMEM_CTRL_BYTES = 6144;     // barriers + tmem ptr + row norms + merge area
constexpr int SMEM_LIMIT = 232448;        // 227 KiB opt-in maximum per CTA

struct AssignParams {
  int64_t N;
  int D, K, Kpad, BN;
  int n_a, n_passes;   // passes: 0:(a0,hi) 1:(a0,lo) 2:(a1,hi)
  int KB;              // ceil(D / 64)
  int n_stages, n_xstages;
  int num_row_tiles, num_code_tiles;
  float margin_rel;
  const float* cmax;   // [1]
  int32_t* idx;
  vqb_flag_entry* flagged;
  int32_t* flag_count;
  float* dbg_best;
};

struct RowState {  // running arg-max of one row (slice) + candidates inside the error band
  float best, thr, W;
  int i0, i1, n;
  __device__ __forceinline__ void init(float w) { W = w; best = -INFINITY; thr = -INFINITY; i0 = 0; i1 = -1; n = 0; }
  // called only for elements with v > thr
  __device__ __forceinline__ void hit(float v, int c) {
    const bool nb = v > best;
    n = (v - best > W) ? 1 : n + 1;   // a clear new leader drops every earlier candidate out of the band
    i1 = nb ? i0 : c;
    i0 = nb ? c : i0;
    best = fmaxf(best, v);
    thr = best - W;
  }
};

struct MergeSlot { float best; int i0, i1, n; };

struct Ctrl {  // lives at the start of dynamic smem
  uint64_t a_full[MAX_A_SUB], a_empty[MAX_A_SUB];
  uint64_t a_read;                       // epilogue finished reading A (row norms)
  uint64_t b_full[MAX_STAGES], b_empty[MAX_STAGES];
  uint64_t x_full[2], x_empty[2];        // bias blocks
  uint64_t t_full[2], t_empty[2];        // TMEM accumulator stages
  uint32_t tmem_base;
  uint32_t pad;
  float xn2[2][BM];                      // row norms, one copy per column-half (no cross-warp sync needed)
  MergeSlot merge[2][BM];                // slice states of the upper column-half warps, double buffered
};
static_assert(sizeof(Ctrl) <= SMEM_CTRL_BYTES, "control block too large");

// 32-byte-swizzle K-major descriptor (the [rows][16 bf16] bias operands): 8-row groups are 256 B apart.
__device__ __forceinline__ uint64_t umma_smem_desc_sw32(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;  // SWIZZLE_32B
  return d;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
vq_assign_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmX, const AssignParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = (smem_base + SMEM_CTRL_BYTES + 1023u) & ~1023u;    // swizzled tiles need 1024 B alignment
  const uint8_t* a_gen = smem + (a_base - smem_base);                         // same place, generic address
  const int n_sub = p.n_a * p.KB;
  const uint32_t aext_base = a_base + n_sub * A_SUB_BYTES;                    // 4 KiB
  const uint32_t b_stage_bytes = p.BN * BK * 2;
  const uint32_t x_stage_bytes = p.BN * 32;
  const uint32_t xb_base = aext_base + AEXT_BYTES;                            // n_xstages * BN*32
  const uint32_t b_base = (xb_base + p.n_xstages * x_stage_bytes + 1023u) & ~1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ------------------------------------------------------------------ one-time setup
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < n_sub; ++s) {
      mbar_init(smem_u32(&ctrl->a_full[s]), 1);
      mbar_init(smem_u32(&ctrl->a_empty[s]), 1);
    }
    mbar_init(smem_u32(&ctrl->a_read), NUM_EPI_WARPS);
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(smem_u32(&ctrl->b_full[s]), 1);
      mbar_init(smem_u32(&ctrl->b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&ctrl->x_full[s]), 1);
      mbar_init(smem_u32(&ctrl->x_empty[s]), 1);
      mbar_init(smem_u32(&ctrl->t_full[s]), 1);
      mbar_init(smem_u32(&ctrl->t_empty[s]), NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&ctrl->tmem_base), TMEM_COLS);
    tmem_relinquish();
  }
  if (threadIdx.x < BM) {
    // constant A-side bias operand: row r = [1 1 1 0 ... 0] (16 bf16 = two 16-byte chunks), 32-byte swizzle:
    // chunk j of row r lives at r*32 + ((j ^ ((r >> 2) & 1)) << 4)
    const int r = threadIdx.x;
    uint8_t* row = const_cast<uint8_t*>(a_gen) + (aext_base - a_base) + r * 32;
    const int sw = (r >> 2) & 1;
    *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
    *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  const int my_tiles = (p.num_row_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  // plane used by pass ps: A plane = (ps == 2), B plane = (ps == 1)
  const int last_pass_a0 = p.n_passes >= 2 ? 1 : 0;  // last pass that reads A plane 0

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      uint32_t it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int tile = blockIdx.x + t * gridDim.x;
        const int row0 = tile * BM;
        if (t > 0) mbar_wait(smem_u32(&ctrl->a_read), (t - 1) & 1);  // norms of the previous tile were read
        for (int ct = 0; ct < p.num_code_tiles; ++ct, ++it) {
          {  // bias block of this code tile
            const uint32_t xs = p.n_xstages == 2 ? (it & 1) : 0;
            const uint32_t xph = p.n_xstages == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(smem_u32(&ctrl->x_empty[xs]), xph ^ 1);
            mbar_arrive_expect_tx(smem_u32(&ctrl->x_full[xs]), x_stage_bytes);
            tma_load_3d(xb_base + xs * x_stage_bytes, &tmX, smem_u32(&ctrl->x_full[xs]), 0, ct * p.BN, 0);
          }
          for (int ps = 0; ps < p.n_passes; ++ps) {
            const int bplane = (ps == 1) ? 1 : 0;
            const int aplane = (ps == 2) ? 1 : 0;
            const bool first_use = (ct == 0) && (ps == 0 || ps == 2);
            for (int kb = 0; kb < p.KB; ++kb) {
              if (first_use) {  // refill this A sub-tile as soon as the previous row tile released it
                const int sub = aplane * p.KB + kb;
                mbar_wait(smem_u32(&ctrl->a_empty[sub]), (t & 1) ^ 1);
                mbar_arrive_expect_tx(smem_u32(&ctrl->a_full[sub]), A_SUB_BYTES);
                tma_load_3d(a_base + sub * A_SUB_BYTES, &tmA, smem_u32(&ctrl->a_full[sub]), kb * BK, row0, aplane);
              }
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
      const uint64_t aext_desc = umma_smem_desc_sw32(aext_base);
      int stage = 0;
      uint32_t ph = 0;
      uint32_t it = 0;  // accumulator iteration counter (across row tiles)
      for (int t = 0; t < my_tiles; ++t) {
        for (int ct = 0; ct < p.num_code_tiles; ++ct, ++it) {
          const uint32_t as = it & 1;
          mbar_wait(smem_u32(&ctrl->t_empty[as]), ((it >> 1) & 1) ^ 1);
          const uint32_t d_tmem = tmem_base + as * 256;
          {  // seed the accumulator with -bias
            const uint32_t xs = p.n_xstages == 2 ? (it & 1) : 0;
            const uint32_t xph = p.n_xstages == 2 ? ((it >> 1) & 1) : (it & 1);
            mbar_wait(smem_u32(&ctrl->x_full[xs]), xph);
            tc_fence_after();
            umma_bf16_ss(d_tmem, aext_desc, umma_smem_desc_sw32(xb_base + xs * x_stage_bytes), idesc, 0u);
            umma_commit(smem_u32(&ctrl->x_empty[xs]));
          }
          for (int ps = 0; ps < p.n_passes; ++ps) {
            const int aplane = (ps == 2) ? 1 : 0;
            const bool last_use = (ct == p.num_code_tiles - 1) && (aplane == 1 ? ps == 2 : ps == last_pass_a0);
            for (int kb = 0; kb < p.KB; ++kb) {
              const int sub = aplane * p.KB + kb;
              if (ct == 0) mbar_wait(smem_u32(&ctrl->a_full[sub]), t & 1);
              mbar_wait(smem_u32(&ctrl->b_full[stage]), ph);
              tc_fence_after();
              const uint32_t a_addr = a_base + sub * A_SUB_BYTES;
              const uint32_t b_addr = b_base + stage * b_stage_bytes;
              const int rem = p.D - kb * BK;
              const int ksteps = rem >= BK ? (BK / UMMA_K) : (rem + UMMA_K - 1) / UMMA_K;
              for (int k = 0; k < ksteps; ++k)
                umma_bf16_ss(d_tmem, umma_smem_desc_sw128(a_addr + k * UMMA_K * 2),
                             umma_smem_desc_sw128(b_addr + k * UMMA_K * 2), idesc, 1u);
              umma_commit(smem_u32(&ctrl->b_empty[stage]));              // B stage reusable once these MMAs retire
              if (last_use) umma_commit(smem_u32(&ctrl->a_empty[sub])); // ... and this A sub-tile too
              if (++stage == p.n_stages) { stage = 0; ph ^= 1; }
            }
          }
          umma_commit(smem_u32(&ctrl->t_full[as]));  // accumulator complete -> epilogue
        }
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..9)
    const int ew = warp - 2;                 // 0..7
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int half = ew >> 2;                // column half: chunk parity handled by this warp
    const int row_in_tile = lg * 32 + lane;  // TMEM lane == row of the tile
    const int pair_bar = 1 + lg;             // named barrier shared by the two warps of a lane group
    const float cmax = __ldg(p.cmax);
    const int n_chunks = (p.BN + 31) / 32;
    uint32_t it = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      // ---- row norms from the A tile in smem (conflict-free: a warp reads 4 full 128 B rows per request)
      for (int s = 0; s < n_sub; ++s) mbar_wait(smem_u32(&ctrl->a_full[s]), t & 1);
      {
        const int sub = lane >> 3, chunk = lane & 7;
        for (int i = 0; i < 8; ++i) {
          const int r = lg * 32 + i * 4 + sub;
          const uint32_t off = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
          float acc2 = 0.f;
          for (int kb = 0; kb < p.KB; ++kb) {
            float v[8];
            {
              const uint4 u = *reinterpret_cast<const uint4*>(a_gen + kb * A_SUB_BYTES + off);
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[2 * e] = __uint_as_float(w[e] << 16);
                v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
              }
            }
            if (p.n_a == 2) {
              const uint4 u = *reinterpret_cast<const uint4*>(a_gen + (p.KB + kb) * A_SUB_BYTES + off);
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
          if (chunk == 0) ctrl->xn2[half][r] = acc2;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctrl->a_read));
      }
      RowState st;
      st.init(2.f * p.margin_rel * sqrtf(ctrl->xn2[half][row_in_tile]) * cmax + 1e-30f);
      __syncwarp();  // xn2 reads done before this warp rewrites it for the next tile

      for (int ct = 0; ct < p.num_code_tiles; ++ct, ++it) {
        const uint32_t as = it & 1;
        mbar_wait(smem_u32(&ctrl->t_full[as]), (it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + as * 256;
        const int code0 = ct * p.BN;
        for (int ci = half; ci < n_chunks; ci += 2) {
          const int c0 = ci * 32;
          if (p.BN - c0 >= 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_addr + c0, r);
            tmem_wait_ld();
            float m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              m[j] = fmaxf(fmaxf(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])),
                           fmaxf(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
            const float mm = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
            if (mm > st.thr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (m[j] > st.thr) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float v = __uint_as_float(r[4 * j + e]);
                    if (v > st.thr) st.hit(v, code0 + c0 + 4 * j + e);
                  }
                }
              }
            }
          } else {  // BN is a multiple of 16: one 16-column tail
            uint32_t r[16];
            tmem_ld_32x32b_x16(t_addr + c0, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v = __uint_as_float(r[j]);
              if (v > st.thr) st.hit(v, code0 + c0 + j);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctrl->t_empty[as]));
      }

      // ---- merge the two column slices of each row (upper half publishes, lower half finishes the row)
      MergeSlot* slot = &ctrl->merge[t & 1][row_in_tile];
      if (half == 1) {
        slot->best = st.best; slot->i0 = st.i0; slot->i1 = st.i1; slot->n = st.n;
        named_bar_sync(pair_bar, 64);
      } else {
        named_bar_sync(pair_bar, 64);
        const float ob = slot->best;
        const int oi0 = slot->i0, oi1 = slot->i1, on = slot->n;
        // candidates of a slice count only if that slice's best is inside the band of the overall best
        const float best = fmaxf(st.best, ob);
        const bool mine_in = st.best >= best - st.W, other_in = ob >= best - st.W;
        const int n = (mine_in ? st.n : 0) + (other_in ? on : 0);
        int i0, i1;
        if (st.best > ob || (st.best == ob && st.i0 < oi0)) { i0 = st.i0; i1 = (mine_in && st.n >= 2) ? st.i1 : oi0; }
        else { i0 = oi0; i1 = (other_in && on >= 2) ? oi1 : st.i0; }
        const int64_t row = static_cast<int64_t>(tile) * BM + row_in_tile;
        if (row < p.N) {
          p.idx[row] = i0;
          if (p.dbg_best) p.dbg_best[row] = best;
          if (n >= 2) {
            const int s = atomicAdd(p.flag_count, 1);
            vqb_flag_entry e;
            e.row = static_cast<int32_t>(row);
            e.cand0 = i0;
            e.cand1 = i1;
            e.count = n;
            p.flagged[s] = e;
          }
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

// bf16 tensor [planes][rows][cols] (row-major) -> boxes of {box_cols, box_rows, 1}, swizzle span == box row bytes,
// out-of-bounds elements read as zero (ragged N / K / D are handled by the zero fill).
static int make_map(CUtensorMap* m, const void* base, int cols, int64_t rows, int planes, int box_cols, int box_rows,
                    CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return VQB_E_DRIVER;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, static_cast<cuuint64_t>(cols) * 2 * static_cast<cuuint64_t>(rows)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQB_OK : VQB_E_DRIVER;
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_padded_codes(int K) {
  if (K <= 0) return 0;
  const int BN = code_tile(K);
  return (K + BN - 1) / BN * BN;
}

extern "C" int vqb_assign(const void* a_planes, int n_a, int64_t N, int D, const void* b_planes, const void* bext,
                          const float* cmax, int K, float margin_rel, int n_passes, int32_t* idx,
                          vqb_flag_entry* flagged, int32_t* flag_count, float* dbg_best, void* stream) {
  if (!a_planes || !b_planes || !bext || !cmax || !idx || !flagged || !flag_count) return VQB_E_INVALID;
  if (N <= 0 || D <= 0 || K <= 0 || (n_a != 1 && n_a != 2)) return VQB_E_INVALID;
  if (n_passes == 0) n_passes = (n_a == 2) ? 3 : 2;
  if (n_passes < 1 || n_passes > 3 || (n_passes == 3 && n_a != 2)) return VQB_E_INVALID;
  if (D % 8 != 0) return VQB_E_UNSUPPORTED;
  const int KB = (D + BK - 1) / BK;
  if (n_a * KB > MAX_A_SUB) return VQB_E_UNSUPPORTED;
  if (N > (static_cast<int64_t>(1) << 31) - BM) return VQB_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a_planes) | reinterpret_cast<uintptr_t>(b_planes) | reinterpret_cast<uintptr_t>(bext)) & 15)
    return VQB_E_ALIGN;
  int rc = check_device();
  if (rc) return rc;

  AssignParams p;
  p.N = N; p.D = D; p.K = K;
  p.BN = code_tile(K);
  p.Kpad = vqb_padded_codes(K);
  if (n_passes < 3) n_a = 1;  // plane 1 of A is only read by pass 2
  p.n_a = n_a; p.n_passes = n_passes; p.KB = KB;
  p.num_row_tiles = static_cast<int>((N + BM - 1) / BM);
  p.num_code_tiles = p.Kpad / p.BN;
  p.margin_rel = margin_rel;
  p.cmax = cmax; p.idx = idx; p.flagged = flagged; p.flag_count = flag_count; p.dbg_best = dbg_best;
  const int a_bytes = n_a * KB * A_SUB_BYTES;
  const int b_stage = p.BN * BK * 2;
  const int x_stage = p.BN * 32;
  const int fixed = SMEM_CTRL_BYTES + 1024 /*align*/ + a_bytes + AEXT_BYTES + 1024 /*align of B ring*/;
  int xstages = 2;
  int stages = (SMEM_LIMIT - fixed - xstages * x_stage) / b_stage;
  if (stages < 3) {  // tight (fp32 split input, D = 256): single-buffer the bias block to keep B stages
    xstages = 1;
    stages = (SMEM_LIMIT - fixed - xstages * x_stage) / b_stage;
  }
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return VQB_E_UNSUPPORTED;
  p.n_stages = stages;
  p.n_xstages = xstages;
  const int smem_bytes = fixed + xstages * x_stage + stages * b_stage;

  CUtensorMap tmA, tmB, tmX;
  rc = make_map(&tmA, a_planes, D, N, n_a, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);  // plane stride = N*D either way
  if (rc) return rc;
  rc = make_map(&tmB, b_planes, D, p.Kpad, 2, BK, p.BN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_map(&tmX, bext, 16, p.Kpad, 1, 16, p.BN, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vq_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  int grid = p.num_row_tiles < num_sms() ? p.num_row_tiles : num_sms();
  vq_assign_kernel<<<grid, NUM_THREADS, smem_bytes, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, tmX, p);
  return static_cast<int>(cudaGetLastError());
}
