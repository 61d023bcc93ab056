// 3x3 convolution with a shared-memory halo tile: the activation tile (8 x TH output pixels plus a one-pixel
// border, 64 channels) is fetched ONCE per 64-channel chunk and all nine filter taps are read from it through
// shifted UMMA descriptors, instead of one TMA box per tap.  The one-tile-per-tap kernel (conv_gemm.cu) is
// bound by the chip-wide L2->SM bandwidth on the wide, shallow layers (profiles/r01_conv_sweep.txt: ~42 B/clk/SM);
// this variant cuts the activation traffic 4x and (one CTA per SM only) optionally keeps the whole weight
// n-tile resident in shared memory when it fits (N <= BN, e.g. the 64-channel 224x224 layers).
//
// Shared-memory geometry of one activation stage (K-major, 128-byte rows, TMA SWIZZLE_128B): one box
// {64 ch, 10 px, TH+2 rows}; halo pixel (hy, hx) sits at row hy*10 + hx.  For filter tap (r, s) the MMA row group
// g (= output row g of the tile, 8 pixels) starts at ((g + r)*10 + s) * 128 B: stride-byte-offset 1280 and a
// start address that is only 128-byte aligned.  Bring-up on the B200 (profiles/r01_halo_descriptor_probe.txt)
// showed that the 128B swizzle of tcgen05 operands is a function of the absolute shared-memory address, so the
// descriptor keeps base_offset = 0 (base_offset = s gives wrong results).
//
// Two CTAs are kept co-resident per SM (<= 113 KB smem, <= 256 TMEM columns each): the tensor pipe issues one
// M=128 MMA per ~97 cycles from a single CTA's dependent chain regardless of N <= 128, and interleaves the chains
// of two CTAs (profiles/r01_mma_microbench.txt), so N=64/128 layers need the second CTA to fill the pipe.
// Everything else (TMEM multi-buffered accumulators, warp roles, fused epilogues, TMA store) follows the
// persistent kernel in conv_gemm.cu.
#include <cstring>
#include "conv_common.cuh"

namespace ug {

__device__ __forceinline__ uint64_t umma_desc_sw128_ex(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

static constexpr int kHaloPitch = 10;  // 8 output pixels + one halo pixel on each side

struct HaloParams {
  int TH;            // output rows per tile (tile = 8 x TH pixels)
  int a_stage_bytes; // bytes of one activation stage
  int sa, sb;        // activation / weight pipeline depths
  int b_resident;    // whole weight n-tile kept in smem (requires n_tiles == 1, one CTA per SM)
};

template <int kAct>
__global__ void __launch_bounds__(kThreads, 2) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const __grid_constant__ CUtensorMap tmO,
                                                                   const ConvKParams p, const HaloParams hp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_tile_bytes = p.BN * 128;
  const int obuf_bytes = p.tma_store ? kABytesPerStage : 0;  // one 64-channel sub-tile per staging buffer
  const int nb_tiles = hp.b_resident ? 9 * p.kchunks : hp.sb;  // weight tiles held in smem
  uint8_t* sA = smem;
  uint8_t* sB = sA + hp.sa * hp.a_stage_bytes;
  uint8_t* sO = sB + nb_tiles * b_tile_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sO + p.obufs * obuf_bytes);
  uint64_t* a_empty = a_full + hp.sa;
  uint64_t* b_full = a_empty + hp.sa;   // hp.sb entries (entry 0 only when resident)
  uint64_t* b_empty = b_full + hp.sb;
  uint64_t* acc_full = b_empty + hp.sb;
  uint64_t* acc_empty = acc_full + p.acc_stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + p.acc_stages);
  float* sScale = reinterpret_cast<float*>(tmem_ptr + 2);
  float* sBias = sScale + p.npad;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.tma_store) prefetch_tmap(&tmO);
    for (int i = 0; i < hp.sa; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < hp.sb; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < p.acc_stages; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.npad; i += kThreads) {
    sScale[i] = (i < p.N) ? (p.scale ? p.scale[i] : 1.0f) : 0.0f;
    sBias[i] = (i < p.N && p.bias) ? p.bias[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (hp.b_resident) {  // whole weight n-tile, once
        mbar_arrive_expect_tx(&b_full[0], (uint32_t)(9 * p.kchunks * b_tile_bytes));
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int tap = 0; tap < 9; ++tap)
            tma_load_2d(sB + (kc * 9 + tap) * b_tile_bytes, &tmB, &b_full[0], (tap * p.kchunks + kc) * 64, 0);
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const uint32_t a_tx = (uint32_t)(kHaloPitch * (hp.TH + 2) * 128);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t % p.m_tiles, nt = t / p.m_tiles;
        const int x0 = (mt % p.tiles_x) * 8;
        const int y0 = ((mt / p.tiles_x) % p.tiles_y) * hp.TH;
        const int n = mt / (p.tiles_x * p.tiles_y);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_empty[as], aph ^ 1);
          mbar_arrive_expect_tx(&a_full[as], a_tx);
          tma_load_4d(sA + as * hp.a_stage_bytes, &tmA, &a_full[as], kc * 64, x0 - 1, y0 - 1, n);
          if (++as == hp.sa) {
            as = 0;
            aph ^= 1;
          }
          if (!hp.b_resident) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              mbar_arrive_expect_tx(&b_full[bs], (uint32_t)b_tile_bytes);
              tma_load_2d(sB + bs * b_tile_bytes, &tmB, &b_full[bs], (tap * p.kchunks + kc) * 64, nt * p.BN);
              if (++bs == hp.sb) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN);
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_phase = 0;
      if (hp.b_resident) {
        mbar_wait(&b_full[0], 0);
        tc_fence_after();
      }
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.BN;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + as * hp.a_stage_bytes);
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, s = tap - r * 3;
            uint32_t b_addr;
            if (hp.b_resident) {
              b_addr = smem_u32(sB + (kc * 9 + tap) * b_tile_bytes);
            } else {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              b_addr = smem_u32(sB + bs * b_tile_bytes);
            }
            const uint64_t ad = umma_desc_sw128_ex(a_base + (r * kHaloPitch + s) * 128, kHaloPitch * 128, 0);
            const uint64_t bd = umma_desc_sw128(b_addr);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
            if (!hp.b_resident) {
              umma_commit(&b_empty[bs]);
              if (++bs == hp.sb) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          umma_commit(&a_empty[as]);
          if (++as == hp.sa) {
            as = 0;
            aph ^= 1;
          }
        }
        umma_commit(&acc_full[acc]);
        if (++acc == p.acc_stages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps, 1 pixel / thread)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int etid = threadIdx.x - 64;
    const int tx = row & 7;
    const int ty = row >> 3;
    const bool row_in_tile = ty < hp.TH;
    int acc = 0, obuf = 0;
    uint32_t acc_phase = 0;

    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mt = t % p.m_tiles, nt = t / p.m_tiles;
      const int x0 = (mt % p.tiles_x) * 8;
      const int y0 = ((mt / p.tiles_x) % p.tiles_y) * hp.TH;
      const int n = mt / (p.tiles_x * p.tiles_y);
      const int ncol0 = nt * p.BN;
      const int x = x0 + tx, y = y0 + ty;
      const bool valid = row_in_tile && (x < p.W) && (y < p.H);
      const long long pix = (long long)y * p.W + x;
      const __nv_bfloat16* add_row =
          reinterpret_cast<const __nv_bfloat16*>(p.add) + (long long)n * p.add_bstride + pix * p.add_cstride + ncol0;
      const float* gate_row = p.gate + (long long)n * p.N + ncol0;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.BN;
      float dot = 0.0f;
      const int ncols = min(p.BN, p.N - ncol0);

      uint32_t v[16];
      __syncwarp();
      tmem_ld16(taddr, v);
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        const bool sub_start = (c0 & 63) == 0;
        if (p.tma_store && sub_start) {
          // staging buffer `obuf` must no longer be read by the TMA store issued obufs sub-tiles ago
          if (etid == 0) {
            if (p.obufs == 2) bulk_wait_group_read<1>();
            else bulk_wait_group_read<0>();
          }
          named_bar_sync(1, 128);
        }
        tmem_ld_wait();
        float f[16];
        epi_math16<kAct>(v, f, sScale, sBias, ncol0 + c0);
        __syncwarp();
        if (c0 + 16 < ncols) tmem_ld16(taddr + c0 + 16, v);
        if (p.mode == UG_EPI_OUTC) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dot += f[j] * __ldg(p.outc_w + ncol0 + c0 + j);
          continue;
        }
        const int groups = (c0 + 16 <= ncols) ? 2 : 1;
        if ((p.mode == UG_EPI_ADD || p.mode == UG_EPI_GATE) && valid) {
          for (int g = 0; g < groups; ++g) epi_add_gate8(p, f + g * 8, add_row + c0 + g * 8, gate_row + c0 + g * 8);
        }
        uint8_t* so_row = sO + obuf * obuf_bytes + row * 128;
        for (int g = 0; g < groups; ++g) {
          uint4 o;
          o.x = pack_bf16x2(f[g * 8 + 0], f[g * 8 + 1]);
          o.y = pack_bf16x2(f[g * 8 + 2], f[g * 8 + 3]);
          o.z = pack_bf16x2(f[g * 8 + 4], f[g * 8 + 5]);
          o.w = pack_bf16x2(f[g * 8 + 6], f[g * 8 + 7]);
          const int chunk = ((c0 & 63) >> 3) + g;
          *reinterpret_cast<uint4*>(so_row + ((chunk ^ (row & 7)) << 4)) = o;
        }
        const bool sub_end = ((c0 + 16) & 63) == 0 || c0 + 16 >= ncols;
        if (sub_end) {
          if (c0 + 16 >= ncols) {  // all TMEM reads of this accumulator are done: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (etid == 0) {
            tma_store_4d(&tmO, sO + obuf * obuf_bytes, ncol0 + (c0 & ~63), x0, y0, n);
            bulk_commit_group();
          }
          if (p.obufs == 2) obuf ^= 1;
        }
      }
      if (p.mode == UG_EPI_OUTC) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (valid) {
          const float logit = dot + p.outc_b;
          const long long o = ((long long)n * p.H + y) * p.W + x;
          p.logits[o] = logit;
          const float sg = 1.0f / (1.0f + expf(-logit));  // torch.sigmoid(seg_out) > 0.5 in fp32
          p.mask[o] = sg > 0.5f ? 1 : 0;
        }
      }
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.tma_store && etid == 0) bulk_wait_group_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

static inline int cdiv_i(int a, int b) { return (a + b - 1) / b; }

// Fills L for the halo variant with `ctas_per_sm` (1 or 2) co-resident CTAs.  Returns UG_EUNSUPPORTED when the
// shape does not fit this kernel.
int conv_halo_prepare(ug_engine* h, const ug_conv_desc* d, int BN, int ctas_per_sm, ConvLaunch* L) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return set_error(h, UG_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (d->R != 3 || d->S != 3 || d->pad != 1 || d->up == 2)
    return set_error(h, UG_EUNSUPPORTED, "conv(halo): 3x3 pad-1 stride-1 convolutions only");
  const int TH = cdiv_i(d->H, cdiv_i(d->H, 16));  // <= 16 rows per tile, no wasted tile rows
  const int cin_pad = cdiv_i(d->Cin, 64) * 64;
  const int kchunks = cin_pad / 64;
  const int n_tiles = cdiv_i(d->N, BN);
  const int npad = n_tiles * BN;
  const long long ktot = 9LL * cin_pad;
  const int tma_store = d->mode != UG_EPI_OUTC;
  const int obuf_bytes = tma_store ? kABytesPerStage : 0;  // staging is per 64-channel sub-tile
  const int acc_stages = std::max(1, std::min(4, (512 / ctas_per_sm) / BN));
  if (acc_stages * BN * ctas_per_sm > 512) return set_error(h, UG_EUNSUPPORTED, "conv(halo): BN too large for %d CTAs/SM", ctas_per_sm);
  const int a_stage = ((kHaloPitch * (TH + 2) * 128 + 1023) / 1024) * 1024;
  const int b_tile = BN * 128;

  HaloParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.TH = TH;
  hp.a_stage_bytes = a_stage;
  const int fixed = 1024 + 8 * (2 * 4 + 2 * 12 + 2 * acc_stages) + 16 + 2 * npad * (int)sizeof(float);
  const int budget = (ctas_per_sm == 2 ? 113 * 1024 : 227 * 1024) - fixed;
  int obufs = tma_store ? 2 : 0;
  const long long resB = 9LL * kchunks * b_tile;
  if (ctas_per_sm == 1 && n_tiles == 1 && resB + 2 * a_stage + (tma_store ? obuf_bytes : 0) <= budget) {
    hp.b_resident = 1;
    hp.sb = 1;
    if (resB + 2 * a_stage + obufs * obuf_bytes > budget) obufs = 1;
    hp.sa = (int)std::min<long long>(4, (budget - resB - obufs * obuf_bytes) / a_stage);
  } else {
    hp.b_resident = 0;
    hp.sa = 2;
    long long rest = budget - 2LL * a_stage - obufs * obuf_bytes;
    if (rest < 4LL * b_tile && obufs == 2) {
      obufs = 1;
      rest = budget - 2LL * a_stage - obuf_bytes;
    }
    hp.sb = (int)std::min<long long>(12, rest / b_tile);
    if (hp.sb < 3) return set_error(h, UG_EUNSUPPORTED, "conv(halo): tile does not fit in shared memory");
    if (hp.sb >= 9 && rest - 9LL * b_tile >= a_stage) {  // room for a third activation stage
      hp.sa = 3;
      hp.sb = (int)std::min<long long>(12, (rest - a_stage) / b_tile);
    }
  }

  ConvKParams& p = L->p;
  memset(&p, 0, sizeof(p));
  p.H = d->H; p.W = d->W; p.B = d->B;
  p.TW = 8; p.TH = TH; p.TN = 1;
  p.tiles_x = cdiv_i(d->W, 8);
  p.tiles_y = cdiv_i(d->H, TH);
  p.R = 3; p.S = 3; p.pad = 1;
  p.kchunks = kchunks; p.num_k = 9 * kchunks;
  p.N = d->N; p.BN = BN; p.stages = hp.sa;
  int tcols = 32;
  while (tcols < acc_stages * BN) tcols <<= 1;
  p.tmem_cols = tcols;
  p.scale = d->scale; p.bias = d->bias;
  p.act = d->act; p.mode = d->mode;
  p.out = d->out; p.out_cstride = d->out_cstride;
  p.up = 1; p.OH = d->H; p.OW = d->W;
  p.add = d->add; p.add_bstride = d->add_bstride; p.add_cstride = d->add_cstride;
  p.gate = d->gate; p.outc_w = d->outc_w; p.outc_b = d->outc_b;
  p.logits = d->logits; p.mask = d->mask;
  p.m_tiles = p.tiles_x * p.tiles_y * d->B; p.n_tiles = n_tiles; p.acc_stages = acc_stages;
  p.tma_store = tma_store; p.obufs = obufs; p.npad = npad;
  L->variant = 3;
  L->halo_mode = ctas_per_sm;
  L->halo_TH = TH; L->halo_a_stage = a_stage; L->halo_copy = 0;
  L->halo_sa = hp.sa; L->halo_sb = hp.sb; L->halo_bres = hp.b_resident;

  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_cstride * 2, (cuuint64_t)d->W * d->in_cstride * 2,
                             (cuuint64_t)d->H * d->W * d->in_cstride * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kHaloPitch, (cuuint32_t)(TH + 2), 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&L->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->in), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(h, UG_ECUDA, "conv(halo): activation tensor map encode failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)npad};
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&L->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(h, UG_ECUDA, "conv(halo): weight tensor map encode failed (%d)", (int)r);
  }
  if (tma_store) {
    cuuint64_t dims[4] = {(cuuint64_t)d->N, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->out_cstride * 2, (cuuint64_t)d->W * d->out_cstride * 2,
                             (cuuint64_t)d->H * d->W * d->out_cstride * 2};
    cuuint32_t box[4] = {64, 8, (cuuint32_t)TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&L->tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d->out, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(h, UG_ECUDA, "conv(halo): output tensor map encode failed (%d)", (int)r);
  } else {
    memset(&L->tmO, 0, sizeof(L->tmO));
  }
  const long long total = (long long)p.m_tiles * n_tiles;
  L->grid = dim3((unsigned)std::min<long long>(total, (long long)h->num_sms * ctas_per_sm), 1, 1);
  const int nb_tiles = hp.b_resident ? 9 * kchunks : hp.sb;
  L->smem = 1024 + (size_t)hp.sa * a_stage + (size_t)nb_tiles * b_tile + (size_t)obufs * obuf_bytes +
            8 * (2 * hp.sa + 2 * hp.sb + 2 * acc_stages) + 16 + 2 * (size_t)npad * sizeof(float);
  if (L->smem > (size_t)(ctas_per_sm == 2 ? 113 : 227) * 1024)
    return set_error(h, UG_EUNSUPPORTED, "conv(halo): shared memory request %zu too large", L->smem);
  return UG_OK;
}

int conv_halo_launch(ug_engine* h, const ConvLaunch* L, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
    const void* fns[] = {(const void*)conv3x3_halo_kernel<UG_ACT_NONE>, (const void*)conv3x3_halo_kernel<UG_ACT_RELU>,
                         (const void*)conv3x3_halo_kernel<UG_ACT_GELU>};
    for (const void* f : fns)
      if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return check_cuda(h, e, "cudaFuncSetAttribute(conv3x3_halo_kernel)");
    attr_set = true;
  }
  HaloParams hp;
  hp.TH = L->halo_TH; hp.a_stage_bytes = L->halo_a_stage;
  hp.sa = L->halo_sa; hp.sb = L->halo_sb; hp.b_resident = L->halo_bres;
  const int act = L->p.act;
  if (act == UG_ACT_RELU)
    conv3x3_halo_kernel<UG_ACT_RELU><<<L->grid, kThreads, L->smem, s>>>(L->tmA, L->tmB, L->tmO, L->p, hp);
  else if (act == UG_ACT_GELU)
    conv3x3_halo_kernel<UG_ACT_GELU><<<L->grid, kThreads, L->smem, s>>>(L->tmA, L->tmB, L->tmO, L->p, hp);
  else
    conv3x3_halo_kernel<UG_ACT_NONE><<<L->grid, kThreads, L->smem, s>>>(L->tmA, L->tmB, L->tmO, L->p, hp);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "conv3x3_halo_kernel launch");
}

}  // namespace ug
