// raster.cuh -- K2: tile-based first-person rasteriser, one thread block per environment.
//
// Replaces MiniWorldEnv.render_obs / _render_world / FrameBuffer.resolve / get_depth_map of
// the reference (miniworld.py:1064-1086, 1177-1236; opengl.py:339-435), i.e. the whole
// OpenGL draw + MSAA resolve + glReadPixels round trip, for N environments per launch.
//
// Structure of one block (env i, 10 warps):
//   A. thread 0 derives the camera (raster_core.cuh: make_camera) and the frame's draw list.
//   B. geometry: one thread per draw item (static room quad or box face) transforms, lights
//      and sets up <= 2 triangles; survivors of frustum / back-face culling are compacted IN
//      DRAW ORDER into shared memory (block-wide ballot/prefix scan) -- the set-up triangles
//      of a frame never touch HBM.
//   C. raster: one warp per 8x8 pixel tile (lane = column x, rows y and y+4).  Per chunk of
//      32 triangles every lane tests one triangle's bbox / edge functions against the tile
//      and a warp ballot yields the tile's coverage list; hits are applied in order to the
//      per-sample (depth16, triangle) keys held in registers.
//   D. resolve: each pixel shades the distinct triangles its samples see (perspective-
//      correct Gouraud x trilinear texture), box-filters, converts to unorm8; the tile is
//      transposed through shared memory and written as 8-byte row segments; depth (sample 0's
//      16-bit code -> metres) goes out as 32-byte row segments.
// HBM traffic per env-step is the framebuffer written once (+ L2-resident template reads).
#pragma once
#include "raster_core.cuh"

#ifdef __CUDACC__

#define MWB_RENDER_THREADS 320
#define MWB_RENDER_WARPS (MWB_RENDER_THREADS / 32)

struct SmemTris {
  const TriRec* t;
  __device__ const TriRec& operator()(uint32_t slot) const { return t[slot]; }
};

template <int MSAA>
__global__ void __launch_bounds__(MWB_RENDER_THREADS)
render_kernel(DevState S, RenderAssets A, uint8_t* __restrict__ obs, float* __restrict__ depth, int tri_cap,
              int* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TriRec* tris = reinterpret_cast<TriRec*>(smem_raw);
  __shared__ Camera cam;
  __shared__ ItemMap imap;
  __shared__ int warp_tot[MWB_RENDER_WARPS];
  __shared__ __align__(8) uint8_t stage[MWB_RENDER_WARPS][8][24];

  const int i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = S.obs_w, H = S.obs_h;

  if (tid == 0) {
    cam = make_camera(S, i);
    imap = build_item_map(S, i);
  }
  __syncthreads();

  // ---- B. geometry -> shared-memory triangle list, draw order preserved
  int ntris = 0;
  for (int start = 0; start < imap.n_items; start += MWB_RENDER_THREADS) {
    const int idx = start + tid;
    TriRec loc[2];
    int cnt = 0;
    if (idx < imap.n_items) {
      Item it;
      fetch_item(S, A, i, imap, idx, it);
      cnt = item_triangles(cam, it, W, H, loc);
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < MWB_RENDER_WARPS; ++w) {
      int v = warp_tot[w];
      if (w < warp) woff += v;
      total += v;
    }
    const int pos = ntris + woff + incl - cnt;
    for (int k = 0; k < cnt; ++k)
      if (pos + k < tri_cap) tris[pos + k] = loc[k];
    ntris += total;
    __syncthreads();
  }
  if (ntris > tri_cap) {
    if (tid == 0) atomicAdd(overflow, 1);
    ntris = tri_cap;
  }

  // ---- C/D. one warp per 8x8 tile
  const int tiles_x = (W + 7) >> 3, tiles_y = (H + 7) >> 3;
  const int lx = lane & 7, ly = lane >> 3;
  SmemTris fetch{tris};
  for (int tile = warp; tile < tiles_x * tiles_y; tile += MWB_RENDER_WARPS) {
    const int tx0 = (tile % tiles_x) << 3, ty0 = (tile / tiles_x) << 3;
    const int px = tx0 + lx, pya = ty0 + ly, pyb = pya + 4;
    uint32_t ka[MSAA], kb[MSAA];
#pragma unroll
    for (int s = 0; s < MSAA; ++s) ka[s] = kb[s] = MWB_SKY_KEY;

    for (int cb = 0; cb < ntris; cb += 32) {
      const int j = cb + lane;
      bool hit = false;
      if (j < ntris) {
        const TriRec& t = tris[j];
        const int bx0 = t.bx & 0xFFFF, bx1 = t.bx >> 16, by0 = t.by & 0xFFFF, by1 = t.by >> 16;
        hit = bx0 <= tx0 + 7 && bx1 >= tx0 && by0 <= ty0 + 7 && by1 >= ty0;
        if (hit) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {   // tile entirely outside one edge?
            float cx = t.A[k] > 0.0f ? (float)(tx0 + 8) : (float)tx0;
            float cy = t.B[k] > 0.0f ? (float)(ty0 + 8) : (float)ty0;
            if (t.A[k] * cx + t.B[k] * cy + t.C[k] + t.R[k] < 0.0f) hit = false;
          }
        }
      }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const TriRec& t = tris[cb + b];
        raster_pixel(t, cb + b, px, pya, MSAA, ka);
        raster_pixel(t, cb + b, px, pyb, MSAA, kb);
      }
    }

    uint8_t ca[3], cbv[3];
    resolve_pixel(A, cam, fetch, ka, MSAA, px, pya, ca);
    resolve_pixel(A, cam, fetch, kb, MSAA, px, pyb, cbv);
    if (obs != nullptr) {
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        stage[warp][ly][lx * 3 + c] = ca[c];
        stage[warp][ly + 4][lx * 3 + c] = cbv[c];
      }
      __syncwarp();
      if (lane < 24) {   // 8 rows x 3 segments of 8 bytes
        const int row = lane / 3, seg = lane % 3;
        const int py = ty0 + row;
        if (py < H && tx0 + 8 <= W) {
          uint2 v = *reinterpret_cast<const uint2*>(&stage[warp][row][seg * 8]);
          *reinterpret_cast<uint2*>(obs + ((size_t)i * H + py) * W * 3 + (size_t)tx0 * 3 + seg * 8) = v;
        } else if (py < H) {   // ragged right edge (W not a multiple of 8): byte stores
          for (int q = 0; q < 8; ++q) {
            int bcol = seg * 8 + q;
            if (tx0 + bcol / 3 < W) obs[((size_t)i * H + py) * W * 3 + (size_t)tx0 * 3 + bcol] = stage[warp][row][bcol];
          }
        }
      }
    }
    if (depth != nullptr) {
      if (px < W && pya < H) depth[((size_t)i * H + pya) * W + px] = depth_code_to_metres(ka[0] >> 16);
      if (px < W && pyb < H) depth[((size_t)i * H + pyb) * W + px] = depth_code_to_metres(kb[0] >> 16);
    }
  }
}

#endif  // __CUDACC__
