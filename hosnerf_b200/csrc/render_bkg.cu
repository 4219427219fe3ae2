// Whole background render of one ray batch behind ONE C call (hos_render_bkg_fused): the level loop of
// MipNeRF360.forward (S1 src/model/mipnerf360/model.py:331-461) sequenced on the caller's stream -
//   per level:  resample (dilate + anneal + inverse CDF + s_to_t)  ->  fused IPE + MLP (tcgen05)  ->  alpha composite,
// with every intermediate in a caller-provided workspace.  No kernel of its own except the constant level-0 histogram;
// it exists so that a host in any language renders a batch with one FFI call instead of re-implementing the loop.
#include "common.cuh"

namespace hos {

__global__ void level0_histogram_kernel(float* __restrict__ sdist, float* __restrict__ weights, int N, float lo, float hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  sdist[2 * i] = lo;          // one interval [lo, hi] of weight 1 per ray (model.py:346-349)
  sdist[2 * i + 1] = hi;
  weights[i] = 1.f;
}

struct BkgWorkspace {
  float *sdist0, *w0, *sdist[2], *weights[2], *tdist, *density, *rgb, *direnc, *rowbias;
  size_t bytes;
};

static BkgWorkspace carve(const hos_bkg_config* c, int N, void* base) {
  int smax = 1, vmax = 0;
  for (int l = 0; l < c->n_levels; ++l) {
    smax = c->levels[l].n_samples > smax ? c->levels[l].n_samples : smax;
    vmax = c->levels[l].view_dim > vmax ? c->levels[l].view_dim : vmax;
  }
  const int s_last = c->levels[c->n_levels - 1].n_samples;
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
    off += (floats * sizeof(float) + 255) & ~(size_t)255;
    return p;
  };
  BkgWorkspace w;
  const size_t n = (size_t)N;
  w.sdist0 = take(n * 2);
  w.w0 = take(n);
  for (int i = 0; i < 2; ++i) { w.sdist[i] = take(n * (smax + 1)); w.weights[i] = take(n * smax); }
  w.tdist = take(n * (smax + 1));
  w.density = take(n * smax);
  w.rgb = take(n * s_last * 3);
  w.direnc = take(n * (3 + 6 * (size_t)c->deg_view));
  w.rowbias = take(n * (size_t)(vmax > 0 ? vmax : 1));
  w.bytes = off;
  return w;
}

static int check_config(const hos_bkg_config* c, const char* who) {
  HOS_REQUIRE(c, "%s: null config", who);
  HOS_REQUIRE(c->n_levels >= 1 && c->n_levels <= HOS_BKG_MAX_LEVELS, "%s: n_levels must be 1..%d", who, HOS_BKG_MAX_LEVELS);
  for (int l = 0; l < c->n_levels; ++l) {
    const hos_bkg_level& L = c->levels[l];
    HOS_REQUIRE(L.mlp && L.u_base, "%s: level %d: null mlp / u_base", who, l);
    HOS_REQUIRE(l > 0 || !L.dilate, "%s: level 0 has nothing to dilate", who);
    HOS_REQUIRE(L.n_samples >= 2 && L.n_samples <= 341, "%s: level %d: n_samples must be 2..341 (3 S + 1 <= 1024 knots)", who, l);
    HOS_REQUIRE((L.view_W == nullptr) == (L.view_dim == 0), "%s: level %d: view_W and view_dim must come together", who, l);
  }
  HOS_REQUIRE(c->levels[c->n_levels - 1].view_W, "%s: the final level needs the view-direction term", who);
  HOS_REQUIRE(c->basis_host, "%s: null basis", who);
  HOS_REQUIRE(c->deg_view >= 0 && c->deg_view <= 16, "%s: bad deg_view", who);
  return HOS_OK;
}

}  // namespace hos

using namespace hos;

extern "C" {

int hos_render_bkg_workspace(const hos_bkg_config* cfg, int N, size_t* bytes) {
  int st = check_config(cfg, "hos_render_bkg_workspace");
  if (st != HOS_OK) return st;
  HOS_REQUIRE(N >= 0 && bytes, "hos_render_bkg_workspace: bad arguments");
  *bytes = carve(cfg, N, nullptr).bytes;
  return HOS_OK;
}

int hos_render_bkg(const hos_bkg_config* cfg, const float* rays_o, const float* rays_d, const float* viewdirs,
                   const float* radii, int N, void* workspace, size_t workspace_bytes, float* rgb_out,
                   float* sdist_out, float* weights_out, void* stream) {
  HOS_ARCH_GUARD();
  int st = check_config(cfg, "hos_render_bkg");
  if (st != HOS_OK) return st;
  HOS_REQUIRE(N >= 0, "hos_render_bkg: bad N");
  if (N == 0) return HOS_OK;
  HOS_REQUIRE(rays_o && rays_d && viewdirs && radii && rgb_out && workspace, "hos_render_bkg: null pointer");
  HOS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hos_render_bkg: workspace must be 256-byte aligned");
  BkgWorkspace w = carve(cfg, N, workspace);
  HOS_REQUIRE(workspace_bytes >= w.bytes, "hos_render_bkg: workspace too small (%zu < %zu bytes)", workspace_bytes, w.bytes);
  cudaStream_t s = (cudaStream_t)stream;

  level0_histogram_kernel<<<(N + 255) / 256, 256, 0, s>>>(w.sdist0, w.w0, N, cfg->dom_lo, cfg->dom_hi);
  HOS_LAUNCH_CHECK();
  const float* sd_in = w.sdist0;
  const float* w_in = w.w0;
  int m_in = 1;
  for (int l = 0; l < cfg->n_levels; ++l) {
    const hos_bkg_level& L = cfg->levels[l];
    const bool last = l == cfg->n_levels - 1;
    const int S = L.n_samples;
    float* sd = (last && sdist_out) ? sdist_out : w.sdist[l & 1];
    float* wt = (last && weights_out) ? weights_out : w.weights[l & 1];
    st = hos_resample_level(sd_in, w_in, N, m_in, L.dilate, L.dilation, cfg->anneal, cfg->resample_padding, L.u_base,
                            L.jitter, L.jitter ? L.jitter_cols : 0, L.max_jitter, S, cfg->dom_lo, cfg->dom_hi, cfg->s_near,
                            cfg->s_far, sd, w.tdist, stream);
    if (st != HOS_OK) return st;
    const float* rowbias = nullptr;
    if (L.view_W) {          // per-ray view-direction term of the view layer (model.py:243-248), bias folded in
      const int de = 3 + 6 * cfg->deg_view;
      st = hos_pos_enc(viewdirs, N, 0, cfg->deg_view, 1, w.direnc, stream);
      if (st != HOS_OK) return st;
      st = hos_linear_f32(w.direnc, de, de, nullptr, 0, 0, L.view_W, L.view_b, N, L.view_dim, 0, w.rowbias, L.view_dim, stream);
      if (st != HOS_OK) return st;
      rowbias = w.rowbias;
    }
    if (cfg->mlp_events) HOS_CUDA(cudaEventRecord((cudaEvent_t)cfg->mlp_events[2 * l], s));
    st = hos_mlp_forward_ipe(L.mlp, w.tdist, rays_o, rays_d, radii, cfg->basis_host, N, S, rowbias, S, w.density,
                             L.view_W ? w.rgb : nullptr, stream);
    if (st != HOS_OK) return st;
    if (cfg->mlp_events) HOS_CUDA(cudaEventRecord((cudaEvent_t)cfg->mlp_events[2 * l + 1], s));
    st = hos_composite_mip360(w.density, w.tdist, rays_d, last ? w.rgb : nullptr, N, S, cfg->opaque_background, cfg->bg, wt,
                              last ? rgb_out : nullptr, stream);
    if (st != HOS_OK) return st;
    sd_in = sd;
    w_in = wt;
    m_in = S;
  }
  return HOS_OK;
}

}  // extern "C"
