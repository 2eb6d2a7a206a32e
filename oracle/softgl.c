/*
 * softgl.c -- TEST INFRASTRUCTURE (pixel oracle).  Not part of the product; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * A deliberately plain, immediate-mode CPU restatement of what the reference's OpenGL calls
 * produce for one observation (reference = Farama-Foundation/Miniworld @ c660156):
 *
 *   MiniWorldEnv.render_obs     miniworld.py:1177-1221   clear(sky, depth 1), gluPerspective(fov_y,
 *                                                        W/H, 0.04, 100), gluLookAt(cam_pos, +cam_dir, +Y)
 *   Agent.cam_pos / cam_dir     entity.py:476-503
 *   _render_static / _world     miniworld.py:1019-1086   LIGHT0 positional, ambient + diffuse colour
 *                                                        material, smooth shading, DEPTH_TEST, CULL_FACE
 *   Room._render                miniworld.py:401-434     floor / ceiling GL_POLYGON, walls GL_QUADS
 *   Box.render / drawBox        entity.py:409-432, opengl.py:460-503
 *   MeshEnt.render              entity.py:150-161, objmesh.py:280-292
 *   Texture.load                opengl.py:147-184        RGB8 + mipmaps, LINEAR / LINEAR_MIPMAP_LINEAR, REPEAT
 *   FrameBuffer                 opengl.py:197-435        N-sample RGBA32F + DEPTH16, resolve, get_depth_map
 *   render_top_view             miniworld.py:1088-1175   glOrtho map view (Scene.view = 1)
 *   get_visible_ents            miniworld.py:1238-1333   GL_ANY_SAMPLES_PASSED queries (Scene.tri_query)
 *
 * PARITY STATUS: **unpinned**.  The reference ships no golden images and its pixels come out
 * of a third-party GL driver (pyglet<2 -> libGL; Mesa llvmpipe in its CI); neither exists in
 * this image, so this restatement follows the OpenGL 2.1 fixed-function rules plus the
 * conventions listed in DESIGN.md ("pixel spec": sample positions, quad split, LOD formula,
 * mip filter, rounding).  It is the oracle the CUDA rasteriser is compared against; it is
 * not evidence of what a particular GL driver would have drawn.
 *
 * Structure here is the textbook one (per-sample colour + 16-bit depth buffers, every
 * triangle tested against every sample, resolve at the end) -- nothing is shared with the
 * tile-based kernel in miniworld_b200/csrc except the arithmetic contract: visibility
 * (vertex transform, homogeneous edge functions, z plane, depth code) is float32 with one
 * rounding per operation in the documented order.  Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ZNEAR 0.04
#define ZFAR 100.0
#define MAXLEV 12

typedef struct {
  int w, h, nlev;
  int lw[MAXLEV], lh[MAXLEV];
  uint8_t* lev[MAXLEV]; /* RGB8, row 0 = bottom of the image */
} Tex;

typedef struct {
  int num_tex;
  Tex* tex;
} TexSet;

/* ---- texture set ------------------------------------------------------------------- */

TexSet* softgl_textures_create(int n) {
  TexSet* ts = (TexSet*)calloc(1, sizeof(TexSet));
  ts->num_tex = n;
  ts->tex = (Tex*)calloc((size_t)n, sizeof(Tex));
  return ts;
}

/* texels: RGB8, top row first (PNG order).  pyglet hands GL the rows bottom-up
 * (opengl.py:158-170), so v = 0 is the bottom of the picture.  glGenerateMipmap: 2x2 box
 * filter on the 8-bit values with round-to-nearest ((sum + 2) >> 2); odd sizes halve with
 * floor, the second tap clamped to the last row / column. */
void softgl_textures_set(TexSet* ts, int idx, int w, int h, const uint8_t* texels) {
  Tex* t = &ts->tex[idx];
  t->w = w;
  t->h = h;
  uint8_t* cur = (uint8_t*)malloc((size_t)w * h * 3);
  for (int y = 0; y < h; ++y) memcpy(cur + (size_t)y * w * 3, texels + (size_t)(h - 1 - y) * w * 3, (size_t)w * 3);
  int lev = 0;
  for (;;) {
    t->lw[lev] = w;
    t->lh[lev] = h;
    t->lev[lev] = cur;
    ++lev;
    if ((w == 1 && h == 1) || lev == MAXLEV) break;
    int nw = w > 1 ? w / 2 : 1, nh = h > 1 ? h / 2 : 1;
    uint8_t* nxt = (uint8_t*)malloc((size_t)nw * nh * 3);
    for (int y = 0; y < nh; ++y)
      for (int x = 0; x < nw; ++x) {
        int xa = w > 1 ? 2 * x : 0, xb = w > 1 ? (2 * x + 1 < w ? 2 * x + 1 : w - 1) : 0;
        int ya = h > 1 ? 2 * y : 0, yb = h > 1 ? (2 * y + 1 < h ? 2 * y + 1 : h - 1) : 0;
        for (int c = 0; c < 3; ++c) {
          int s = cur[((size_t)ya * w + xa) * 3 + c] + cur[((size_t)ya * w + xb) * 3 + c] +
                  cur[((size_t)yb * w + xa) * 3 + c] + cur[((size_t)yb * w + xb) * 3 + c];
          nxt[((size_t)y * nw + x) * 3 + c] = (uint8_t)((s + 2) >> 2);
        }
      }
    cur = nxt;
    w = nw;
    h = nh;
  }
  t->nlev = lev;
}

void softgl_textures_destroy(TexSet* ts) {
  if (!ts) return;
  for (int i = 0; i < ts->num_tex; ++i)
    for (int l = 0; l < ts->tex[i].nlev; ++l) free(ts->tex[i].lev[l]);
  free(ts->tex);
  free(ts);
}

/* ---- scene description (filled by the Python harness from a host-side world) ------- */

typedef struct {
  /* agent pose and camera parameters, as float64 (reference entity.py:455-516) */
  double pos[3], dir, cam_height, cam_fwd_disp, cam_pitch, cam_fov_y;
  double sky[3], light_pos[3], light_color[3], light_ambient[3];
  int width, height, samples;
  /* draw list, in GL submission order: triangles with per-vertex position / normal /
   * texcoord and one material colour; tex < 0 = untextured */
  int num_tris;
  const float* tri_pos; /* [T][3][3] world space */
  const float* tri_nrm; /* [T][3][3] */
  const float* tri_uv;  /* [T][3][2] */
  const float* tri_rgb; /* [T][3][3] material (glColor / c3f) per vertex */
  const int* tri_tex;   /* [T] */
  /* view: 0 = the agent's perspective camera; 1 = render_top_view: glOrtho(ortho[0..3] = l, r, b, t,
   * near -100, far 100) under the model-view (x, y, z) -> (x, -z, y) (miniworld.py:1135-1162) */
  int view;
  double ortho[4];
  /* occlusion queries (get_visible_ents): tri_query[t] = query id of triangle t or -1 (NULL: none);
   * query_out[q] is set to 1 when any sample of a triangle of query q passes the depth test */
  const int* tri_query;
  uint8_t* query_out;
  /* cam_mode 1 = the camera as the reference's GL stream states it (oracle/gl_record.py): gluLookAt(eye, center,
   * up) and gluPerspective(fovy, aspect, znear, zfar) arguments instead of the agent's pose and angles.  The basis
   * follows GLU's algorithm (f = normalise(center - eye), s = normalise(f x up), u = s x f) in float64 and is
   * rounded once to float32, like the angle-derived basis of cam_mode 0. */
  int cam_mode;
  double eye[3], center[3], up[3], fovy, aspect, znear, zfar;
  /* LIGHT0's GL_POSITION w component: 1 = positional light at light_pos, 0 = DIRECTIONAL light whose direction
   * (towards the light) is light_pos.  The reference issues (GLfloat * 4)(*self.light_pos + [1]) with light_pos a
   * numpy array (params.py:45-46 turns the defaults into arrays): the `+ [1]` broadcasts, three components are
   * passed and w stays 0 -- a directional light along light_pos + 1 (miniworld.py:1031). */
  double light_w;
} Scene;

/* D3D standard sample patterns in image space (x right, y down), offsets from the pixel's
 * top-left corner */
static const float PAT1[1][2] = {{0.5f, 0.5f}};
static const float PAT4[4][2] = {{0.375f, 0.125f}, {0.875f, 0.375f}, {0.125f, 0.625f}, {0.625f, 0.875f}};
static const float PAT8[8][2] = {{0.5625f, 0.3125f}, {0.4375f, 0.6875f}, {0.8125f, 0.5625f}, {0.3125f, 0.1875f},
                                 {0.1875f, 0.8125f}, {0.0625f, 0.4375f}, {0.6875f, 0.9375f}, {0.9375f, 0.0625f}};

static const float PAT16[16][2] = {{0.5625f, 0.5625f}, {0.4375f, 0.3125f}, {0.3125f, 0.6250f}, {0.7500f, 0.4375f}, {0.1875f, 0.3750f}, {0.6250f, 0.8125f}, {0.8125f, 0.6875f}, {0.6875f, 0.1875f}, {0.3750f, 0.8750f}, {0.5000f, 0.0625f}, {0.2500f, 0.1250f}, {0.1250f, 0.7500f}, {0.0000f, 0.5000f}, {0.9375f, 0.2500f}, {0.8750f, 0.9375f}, {0.0625f, 0.0000f}};

typedef struct {
  float X, Y, W, Z; /* window-homogeneous position: X/W column, Y/W row, Z/W window depth */
  float cx, cy, cz; /* clip coordinates */
  float col[3], u, v;
} Vtx;

static void bilerp(const Tex* t, int level, float u, float v, float out[3]) {
  int w = t->lw[level], h = t->lh[level];
  const uint8_t* px = t->lev[level];
  float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
  float x0f = floorf(x), y0f = floorf(y);
  float fx = x - x0f, fy = y - y0f;
  int x0 = (int)x0f % w, y0 = (int)y0f % h;
  if (x0 < 0) x0 += w;
  if (y0 < 0) y0 += h;
  int x1 = (x0 + 1) % w, y1 = (y0 + 1) % h;
  for (int c = 0; c < 3; ++c) {
    float a = px[((size_t)y0 * w + x0) * 3 + c], b = px[((size_t)y0 * w + x1) * 3 + c];
    float d = px[((size_t)y1 * w + x0) * 3 + c], e = px[((size_t)y1 * w + x1) * 3 + c];
    float top = a + fx * (b - a), bot = d + fx * (e - d);
    out[c] = (top + fy * (bot - top)) / 255.0f;
  }
}

int softgl_render(const Scene* sc, const TexSet* ts, uint8_t* rgb_out, float* depth_out, uint16_t* code_out) {
  const int W = sc->width, H = sc->height, NS = sc->samples;
  const float(*pat)[2] = NS == 1 ? PAT1 : (NS == 4 ? PAT4 : (NS == 8 ? PAT8 : PAT16));
  if (NS != 1 && NS != 4 && NS != 8 && NS != 16) return -1;

  /* --- camera: cam_pos = pos + fwd_disp * dir_vec + (0, h, 0); cam_dir = X rotated by pitch
   * then heading; gluLookAt basis f, s = f x up (normalised), u = s x f, evaluated in float64
   * and rounded once to float32 */
  double ct = cos(sc->dir), st = sin(sc->dir);
  double phi = sc->cam_pitch * 3.141592653589793 / 180.0;
  double cp = cos(phi), sp = sin(phi);
  float eye[3] = {(float)(sc->pos[0] + sc->cam_fwd_disp * ct), (float)(sc->pos[1] + sc->cam_height),
                  (float)(sc->pos[2] - sc->cam_fwd_disp * st)};
  float Sv[3] = {(float)st, 0.0f, (float)ct};
  float Uv[3] = {(float)(-(ct * sp)), (float)cp, (float)(st * sp)};
  float Fv[3] = {(float)(cp * ct), (float)sp, (float)(-(cp * st))};
  double half = sc->cam_fov_y * 3.141592653589793 / 360.0;
  double cot = cos(half) / sin(half);
  float Py = (float)cot, Px = (float)(cot / ((double)W / (double)H));
  float Za = (float)((ZFAR + ZNEAR) / (ZFAR - ZNEAR)), Zb = (float)(2.0 * ZFAR * ZNEAR / (ZFAR - ZNEAR));
  if (sc->cam_mode == 1) {
    double f[3] = {sc->center[0] - sc->eye[0], sc->center[1] - sc->eye[1], sc->center[2] - sc->eye[2]};
    double fl = sqrt((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
    for (int c = 0; c < 3; ++c) f[c] /= fl;
    double s[3] = {f[1] * sc->up[2] - f[2] * sc->up[1], f[2] * sc->up[0] - f[0] * sc->up[2], f[0] * sc->up[1] - f[1] * sc->up[0]};
    double sl = sqrt((s[0] * s[0] + s[1] * s[1]) + s[2] * s[2]);
    for (int c = 0; c < 3; ++c) s[c] /= sl;
    double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
    for (int c = 0; c < 3; ++c) {
      eye[c] = (float)sc->eye[c];
      Sv[c] = (float)s[c];
      Uv[c] = (float)u[c];
      Fv[c] = (float)f[c];
    }
    double radians = sc->fovy / 2 * 3.141592653589793 / 180; /* gluPerspective */
    double cotg = cos(radians) / sin(radians);
    Py = (float)cotg;
    Px = (float)(cotg / sc->aspect);
    Za = (float)((sc->zfar + sc->znear) / (sc->zfar - sc->znear));
    Zb = (float)(2.0 * sc->zfar * sc->znear / (sc->zfar - sc->znear));
  }
  float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
  /* glOrtho's matrix entries: formed in double, stored as float32 */
  float Osx = 0.0f, Otx = 0.0f, Osy = 0.0f, Oty = 0.0f, Osz = (float)(-2.0 / (100.0 - (-100.0)));
  if (sc->view == 1) {
    double l = sc->ortho[0], r = sc->ortho[1], b = sc->ortho[2], t = sc->ortho[3];
    Osx = (float)(2.0 / (r - l));
    Otx = (float)(-((r + l) / (r - l)));
    Osy = (float)(2.0 / (t - b));
    Oty = (float)(-((t + b) / (t - b)));
  }
  float lpos[3], lamb[3], ldif[3], sky[3];
  for (int c = 0; c < 3; ++c) {
    lpos[c] = (float)sc->light_pos[c];
    lamb[c] = (float)sc->light_ambient[c];
    ldif[c] = (float)sc->light_color[c];
    sky[c] = (float)sc->sky[c];
  }

  /* --- buffers: glClearColor(sky), glClearDepth(1.0) */
  size_t nsamp = (size_t)W * H * NS;
  float* cbuf = (float*)malloc(nsamp * 3 * sizeof(float));
  uint16_t* zbuf = (uint16_t*)malloc(nsamp * sizeof(uint16_t));
  for (size_t i = 0; i < nsamp; ++i) {
    cbuf[i * 3 + 0] = sky[0];
    cbuf[i * 3 + 1] = sky[1];
    cbuf[i * 3 + 2] = sky[2];
    zbuf[i] = 65535;
  }

  for (int ti = 0; ti < sc->num_tris; ++ti) {
    Vtx g[3];
    for (int k = 0; k < 3; ++k) {
      const float* p = sc->tri_pos + ((size_t)ti * 3 + k) * 3;
      const float* n = sc->tri_nrm + ((size_t)ti * 3 + k) * 3;
      const float* m = sc->tri_rgb + ((size_t)ti * 3 + k) * 3;
      float we;
      if (sc->view == 1) {
        /* map view: eye = (x, -z, y) exactly (a permutation matrix), clip = glOrtho * eye, w = 1 */
        we = 1.0f;
        g[k].cx = Osx * p[0] + Otx;
        g[k].cy = Osy * (-p[2]) + Oty;
        g[k].cz = Osz * p[1];
      } else {
        /* eye space: subtract the eye, project on (s, u, f) -- each op rounds to float32 */
        float rx = p[0] - eye[0], ry = p[1] - eye[1], rz = p[2] - eye[2];
        float xe = (Sv[0] * rx + Sv[1] * ry) + Sv[2] * rz;
        float ye = (Uv[0] * rx + Uv[1] * ry) + Uv[2] * rz;
        we = (Fv[0] * rx + Fv[1] * ry) + Fv[2] * rz; /* = -z_eye = w_clip */
        g[k].cx = Px * xe;
        g[k].cy = Py * ye;
        g[k].cz = Za * we - Zb;
      }
      g[k].W = we;
      g[k].X = (g[k].cx + we) * hw; /* ((x_ndc + 1) W / 2) w */
      g[k].Y = (we - g[k].cy) * hh; /* ((1 - y_ndc) H / 2) w : row 0 at the top */
      g[k].Z = 0.5f * (g[k].cz + we);
      /* fixed-function lighting, per vertex, normal not renormalised */
      float lx = lpos[0], ly = lpos[1], lz = lpos[2];
      if (sc->light_w != 0.0) { lx -= p[0]; ly -= p[1]; lz -= p[2]; }
      float ll = sqrtf(lx * lx + ly * ly + lz * lz);
      float ndl = (n[0] * lx + n[1] * ly + n[2] * lz) / ll;
      if (ndl < 0.0f) ndl = 0.0f;
      for (int c = 0; c < 3; ++c) {
        float v = m[c] * (0.2f + lamb[c] + ldif[c] * ndl);
        g[k].col[c] = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
      }
      g[k].u = sc->tri_uv[((size_t)ti * 3 + k) * 2 + 0];
      g[k].v = sc->tri_uv[((size_t)ti * 3 + k) * 2 + 1];
    }
    /* trivial frustum rejection in clip space */
    int out_l = 1, out_r = 1, out_b = 1, out_t = 1, out_n = 1, out_f = 1;
    for (int k = 0; k < 3; ++k) {
      out_l &= g[k].cx < -g[k].W;
      out_r &= g[k].cx > g[k].W;
      out_b &= g[k].cy < -g[k].W;
      out_t &= g[k].cy > g[k].W;
      out_n &= g[k].cz < -g[k].W;
      out_f &= g[k].cz > g[k].W;
    }
    if (out_l || out_r || out_b || out_t || out_n || out_f) continue;

    /* In image space (y down) a GL-front (CCW in y-up) triangle is clockwise; work on the
     * vertex order (0, 2, 1) so that front-facing <=> positive determinant, interior E >= 0 */
    const Vtx* v[3] = {&g[0], &g[2], &g[1]};
    float EA[3], EB[3], EC[3];
    for (int k = 0; k < 3; ++k) { /* edge k joins v[k+1] -> v[k+2], opposite v[k] */
      const Vtx* a = v[(k + 1) % 3];
      const Vtx* b = v[(k + 2) % 3];
      EA[k] = a->Y * b->W - a->W * b->Y;
      EB[k] = a->W * b->X - a->X * b->W;
      EC[k] = a->X * b->Y - a->Y * b->X;
    }
    float det = (v[0]->X * EA[0] + v[0]->Y * EB[0]) + v[0]->W * EC[0];
    if (!(det > 0.0f)) continue; /* glCullFace(GL_BACK) + degenerate */
    float ZA = ((v[0]->Z * EA[0] + v[1]->Z * EA[1]) + v[2]->Z * EA[2]) / det;
    float ZB = ((v[0]->Z * EB[0] + v[1]->Z * EB[1]) + v[2]->Z * EB[2]) / det;
    float ZC = ((v[0]->Z * EC[0] + v[1]->Z * EC[1]) + v[2]->Z * EC[2]) / det;
    const Tex* tex = sc->tri_tex[ti] >= 0 ? &ts->tex[sc->tri_tex[ti]] : NULL;

    /* loop bounds only: when the whole triangle is in front of the eye its samples lie
     * inside the bounding box of its projected vertices (padded by one pixel) */
    int bx0 = 0, bx1 = W - 1, by0 = 0, by1 = H - 1;
    if (g[0].W > 1e-3f && g[1].W > 1e-3f && g[2].W > 1e-3f) {
      float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
      for (int k = 0; k < 3; ++k) {
        float sx = g[k].X / g[k].W, sy = g[k].Y / g[k].W;
        xmin = sx < xmin ? sx : xmin; xmax = sx > xmax ? sx : xmax;
        ymin = sy < ymin ? sy : ymin; ymax = sy > ymax ? sy : ymax;
      }
      if (xmax < -1.0f || ymax < -1.0f || xmin > (float)W + 1.0f || ymin > (float)H + 1.0f) continue;
      if (xmin - 1.0f > 0.0f) bx0 = (int)(xmin - 1.0f);
      if (ymin - 1.0f > 0.0f) by0 = (int)(ymin - 1.0f);
      if (xmax + 1.0f < (float)(W - 1)) bx1 = (int)(xmax + 1.0f);
      if (ymax + 1.0f < (float)(H - 1)) by1 = (int)(ymax + 1.0f);
    }
    for (int py = by0; py <= by1; ++py)
      for (int px = bx0; px <= bx1; ++px) {
        unsigned pass = 0;
        uint16_t codes[16];
        for (int s = 0; s < NS; ++s) {
          float xs = (float)px + pat[s][0], ys = (float)py + pat[s][1];
          int inside = 1;
          for (int k = 0; k < 3 && inside; ++k) {
            float e = (EA[k] * xs + EB[k] * ys) + EC[k];
            if (e < 0.0f) inside = 0;
            else if (e == 0.0f) inside = EA[k] > 0.0f || (EA[k] == 0.0f && EB[k] > 0.0f); /* tie rule */
          }
          if (!inside) continue;
          float z = (ZA * xs + ZB * ys) + ZC;
          if (!(z >= 0.0f && z <= 1.0f)) continue; /* near / far clip */
          uint16_t code = (uint16_t)(uint32_t)(z * 65535.0f + 0.5f);
          if (code < zbuf[((size_t)py * W + px) * NS + s]) { /* GL_LESS on DEPTH_COMPONENT16 */
            pass |= 1u << s;
            codes[s] = code;
          }
        }
        if (!pass) continue;
        if (sc->tri_query && sc->query_out && sc->tri_query[ti] >= 0) sc->query_out[sc->tri_query[ti]] = 1;
        /* fragment colour, evaluated once at the pixel centre (multisampling, not supersampling) */
        float cxp = (float)px + 0.5f, cyp = (float)py + 0.5f;
        float e[3], esum = 0.0f;
        for (int k = 0; k < 3; ++k) {
          e[k] = EA[k] * cxp + EB[k] * cyp + EC[k];
          esum += e[k];
        }
        float bw[3] = {e[0] / esum, e[1] / esum, e[2] / esum};
        float col[3];
        for (int c = 0; c < 3; ++c) col[c] = bw[0] * v[0]->col[c] + bw[1] * v[1]->col[c] + bw[2] * v[2]->col[c];
        if (tex) {
          float uu = bw[0] * v[0]->u + bw[1] * v[1]->u + bw[2] * v[2]->u;
          float vv = bw[0] * v[0]->v + bw[1] * v[1]->v + bw[2] * v[2]->v;
          float sa = EA[0] + EA[1] + EA[2], sb = EB[0] + EB[1] + EB[2];
          float ua = v[0]->u * EA[0] + v[1]->u * EA[1] + v[2]->u * EA[2];
          float ub = v[0]->u * EB[0] + v[1]->u * EB[1] + v[2]->u * EB[2];
          float va = v[0]->v * EA[0] + v[1]->v * EA[1] + v[2]->v * EA[2];
          float vb = v[0]->v * EB[0] + v[1]->v * EB[1] + v[2]->v * EB[2];
          float dudx = (ua - uu * sa) / esum * (float)tex->w, dvdx = (va - vv * sa) / esum * (float)tex->h;
          float dudy = (ub - uu * sb) / esum * (float)tex->w, dvdy = (vb - vv * sb) / esum * (float)tex->h;
          float r1 = dudx * dudx + dvdx * dvdx, r2 = dudy * dudy + dvdy * dvdy;
          float rho2 = r1 > r2 ? r1 : r2;
          if (rho2 < 1e-20f) rho2 = 1e-20f;
          float lambda = 0.5f * log2f(rho2);
          float tc[3];
          if (lambda <= 0.0f) {
            bilerp(tex, 0, uu, vv, tc);
          } else if (lambda >= (float)(tex->nlev - 1)) {
            bilerp(tex, tex->nlev - 1, uu, vv, tc);
          } else {
            int l0 = (int)lambda;
            float f = lambda - (float)l0, t0[3], t1[3];
            bilerp(tex, l0, uu, vv, t0);
            bilerp(tex, l0 + 1, uu, vv, t1);
            for (int c = 0; c < 3; ++c) tc[c] = t0[c] + f * (t1[c] - t0[c]);
          }
          for (int c = 0; c < 3; ++c) col[c] *= tc[c]; /* GL_MODULATE */
        }
        for (int s = 0; s < NS; ++s)
          if (pass & (1u << s)) {
            size_t si = ((size_t)py * W + px) * NS + s;
            zbuf[si] = codes[s];
            cbuf[si * 3 + 0] = col[0];
            cbuf[si * 3 + 1] = col[1];
            cbuf[si * 3 + 2] = col[2];
          }
      }
  }

  /* --- resolve: box filter of the samples -> unorm8; depth = sample 0 (GL_NEAREST blit) */
  const float c0 = (float)(-2.0 * ZFAR * ZNEAR), c1 = (float)(ZFAR - ZNEAR), c2 = (float)(ZFAR + ZNEAR);
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      size_t base = ((size_t)py * W + px) * NS;
      for (int c = 0; c < 3; ++c) {
        float acc = 0.0f;
        for (int s = 0; s < NS; ++s) acc += cbuf[(base + s) * 3 + c];
        acc /= (float)NS;
        acc = acc < 0.0f ? 0.0f : (acc > 1.0f ? 1.0f : acc);
        if (rgb_out) rgb_out[((size_t)py * W + px) * 3 + c] = (uint8_t)(int)(acc * 255.0f + 0.5f);
      }
      uint16_t code = zbuf[base];
      if (code_out) code_out[(size_t)py * W + px] = code;
      if (depth_out) { /* FrameBuffer.get_depth_map, float32 arithmetic (opengl.py:427-431) */
        float d = (float)code / 65535.0f;
        float clip = (d - 0.5f) * 2.0f;
        depth_out[(size_t)py * W + px] = c0 / (clip * c1 - c2);
      }
    }
  free(cbuf);
  free(zbuf);
  return 0;
}
