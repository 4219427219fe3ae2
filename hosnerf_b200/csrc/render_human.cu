// Whole human-object render of one ray chunk behind ONE C call (hos_render_human_fused): the body of
// Network._render_rays (S3 core/nets/human_nerf/network.py:427-472, 538-557; S2 :537-556) sequenced on the caller's stream -
//   stratified samples -> LBS warp -> Hann-windowed PE -> non-rigid MLP (tcgen05) -> Fourier PE -> canonical MLP (tcgen05)
//   (-> S2 composite),
// with every intermediate in a caller-provided workspace.  No kernel of its own: it exists so that a host in any language
// renders a chunk with one FFI call (and one CUDA-graph node sequence) instead of re-implementing the chain; the cycle / flow
// side paths only feed the training loss and stay with the caller.
#include "common.cuh"

namespace hos {

struct HumanWorkspace {
  float *z, *pts, *x_skel, *mask, *cnl, *raw;
  unsigned char *pe_nr, *pe_cnl;
  size_t bytes;
};

static size_t tiled_bytes(int64_t rows, int k) {
  return (size_t)((rows + kTileRows - 1) / kTileRows) * (size_t)((k + kTileK - 1) / kTileK) * kTileChunkBytes;
}

static HumanWorkspace carve(const hos_human_config* c, int n, void* base) {
  size_t off = 0;
  auto take = [&](size_t nbytes) {
    void* p = base ? static_cast<void*>(static_cast<char*>(base) + off) : nullptr;
    off += (nbytes + 255) & ~(size_t)255;
    return p;
  };
  HumanWorkspace w;
  const size_t P = (size_t)n * c->n_samples;
  w.z = (float*)take(P * 4);
  w.pts = (float*)take(P * 12);
  w.x_skel = (float*)take(P * 12);
  w.mask = (float*)take(P * 4);
  w.cnl = (float*)take(P * 12);
  w.raw = (float*)take(P * 16);
  w.pe_nr = (unsigned char*)take(tiled_bytes((int64_t)P, 6 * c->nr_freqs));
  w.pe_cnl = (unsigned char*)take(tiled_bytes((int64_t)P, 3 + 6 * c->cnl_freqs));
  w.bytes = off;
  return w;
}

static int check_config(const hos_human_config* c, const char* who) {
  HOS_REQUIRE(c, "%s: null config", who);
  HOS_REQUIRE(c->cnl_mlp && c->t_lin && c->R && c->T && c->vol && c->bbox_min_host && c->bbox_scale_host, "%s: null pointer in config", who);
  HOS_REQUIRE(c->n_samples >= 1 && c->bones >= 1 && c->bones <= 32 && c->grid >= 2, "%s: bad sizes", who);
  HOS_REQUIRE(c->cnl_freqs >= 1 && c->cnl_freqs <= 16 && (!c->nr_mlp || (c->nr_freqs >= 1 && c->nr_freqs <= 16)), "%s: bad frequency counts", who);
  HOS_REQUIRE(!c->nr_mlp || c->hann_w, "%s: the non-rigid MLP needs the Hann window weights", who);
  return HOS_OK;
}

}  // namespace hos

using namespace hos;

extern "C" {

int hos_render_human_workspace(const hos_human_config* cfg, int n, size_t* bytes) {
  int st = check_config(cfg, "hos_render_human_workspace");
  if (st != HOS_OK) return st;
  HOS_REQUIRE(n >= 0 && bytes, "hos_render_human_workspace: bad arguments");
  *bytes = carve(cfg, n, nullptr).bytes;
  return HOS_OK;
}

int hos_render_human(const hos_human_config* cfg, const float* rays_o, const float* rays_d, const float* near, const float* far,
                     int n, void* workspace, size_t workspace_bytes, float* rgb_out, float* acc_out, float* weights_out,
                     float* depth_out, float* raw_out, float* mask_out, float* pts_out, float* z_out, void* stream) {
  HOS_ARCH_GUARD();
  int st = check_config(cfg, "hos_render_human");
  if (st != HOS_OK) return st;
  HOS_REQUIRE(n >= 0, "hos_render_human: bad n");
  if (n == 0) return HOS_OK;
  HOS_REQUIRE(rays_o && rays_d && near && far && workspace, "hos_render_human: null pointer");
  HOS_REQUIRE(!cfg->stage2 || (rgb_out && acc_out && weights_out && depth_out), "hos_render_human: stage 2 needs rgb / acc / weights / depth outputs");
  HOS_REQUIRE(cfg->stage2 || (raw_out && mask_out), "hos_render_human: stage 3 needs raw / mask outputs");
  HOS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hos_render_human: workspace must be 256-byte aligned");
  HumanWorkspace w = carve(cfg, n, workspace);
  HOS_REQUIRE(workspace_bytes >= w.bytes, "hos_render_human: workspace too small (%zu < %zu bytes)", workspace_bytes, w.bytes);
  const int S = cfg->n_samples;
  const int64_t P = (int64_t)n * S;
  float* z = z_out ? z_out : w.z;
  float* pts = pts_out ? pts_out : w.pts;
  float* mask = (!cfg->stage2 && mask_out) ? mask_out : w.mask;
  float* raw = (!cfg->stage2 && raw_out) ? raw_out : w.raw;
  if ((st = hos_human_samples(rays_o, rays_d, near, far, cfg->t_lin, cfg->jitter, n, S, z, pts, stream)) != HOS_OK) return st;
  if ((st = hos_lbs_warp(pts, cfg->R, cfg->T, cfg->vol, cfg->bbox_min_host, cfg->bbox_scale_host, P, cfg->bones, cfg->grid, w.x_skel,
                         mask, stream)) != HOS_OK) return st;
  // Encodings: generated inside the MLP kernel (Fourier prologue) when the handle / batch supports it, else materialised.
  const float* cnl = w.x_skel;
  if (cfg->nr_mlp) {          // x + offset(x)  (network.py:165-172)
    if (cfg->hann_w_host && hos_mlp_fourier_supported(cfg->nr_mlp, P) && cfg->nr_freqs <= 10) {
      st = hos_mlp_forward_fourier(cfg->nr_mlp, w.x_skel, P, cfg->nr_freqs, 0, cfg->hann_w_host, w.x_skel, w.cnl, nullptr, stream);
    } else {
      if ((st = hos_fourier_embed(w.x_skel, P, cfg->nr_freqs, 0, cfg->hann_w, w.pe_nr, 0, 2, stream)) != HOS_OK) return st;
      st = hos_mlp_forward(cfg->nr_mlp, w.pe_nr, P, nullptr, 1, w.x_skel, w.cnl, nullptr, stream);
    }
    if (st != HOS_OK) return st;
    cnl = w.cnl;
  }
  if (cfg->hann_w_host && hos_mlp_fourier_supported(cfg->cnl_mlp, P) && cfg->cnl_freqs <= 10) {
    st = hos_mlp_forward_fourier(cfg->cnl_mlp, cnl, P, cfg->cnl_freqs, 1, nullptr, nullptr, raw, nullptr, stream);
  } else {
    if ((st = hos_fourier_embed(cnl, P, cfg->cnl_freqs, 1, nullptr, w.pe_cnl, 0, 2, stream)) != HOS_OK) return st;
    st = hos_mlp_forward(cfg->cnl_mlp, w.pe_cnl, P, nullptr, 1, nullptr, raw, nullptr, stream);
  }
  if (st != HOS_OK) return st;
  if (cfg->stage2)
    return hos_composite_nerf(raw, mask, z, rays_d, cfg->bgcolor_host, n, S, 1, rgb_out, acc_out, weights_out, depth_out, stream);
  return HOS_OK;
}

}  // extern "C"
