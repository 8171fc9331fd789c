// Fused classification / distillation loss (forward value + gradient wrt the student logits in one
// launch) and the Reasoning output stage.  See include/tvt.h: tvt_distill_loss, tvt_pyramid_head.
// Reference call sites: src/models/transformer.py:35,142 (BCEWithLogits), src/models/frame_transformer.py
// :250-257 (CE on argmax(teacher) + BCE + cosine monitor), src/models/TPN.py:98,112 (sigmoid, mean).
// Latency-bound (B x C <= a few thousand elements): one warp per clip, lanes over classes.
#include "tvt_common.cuh"

namespace tvt {
namespace loss {

struct Params {
  const float* s; const float* t; const float* y; float* losses; float* dl;
  long long B; int C; float w_bce, w_ce, w_kl, T, gscale;
};

__global__ void __launch_bounds__(128) distill_kernel(const Params p) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (row >= p.B) return;
  const float* s = p.s + row * p.C;
  const float* y = p.y + row * p.C;
  const float* t = p.t ? p.t + row * p.C : nullptr;
  const float invBC = 1.0f / (static_cast<float>(p.B) * p.C), invB = 1.0f / static_cast<float>(p.B);
  // pass 1: maxima, BCE, teacher argmax, cosine pieces
  float smax = -INFINITY, tmax = -INFINITY, bce = 0.0f, dot = 0.0f, ns = 0.0f, nt = 0.0f;
  int targ = 0x7fffffff;
  for (int c = lane; c < p.C; c += 32) {
    const float x = s[c];
    smax = fmaxf(smax, x);
    bce += fmaxf(x, 0.0f) - x * y[c] + log1pf(__expf(-fabsf(x)));
    ns += x * x;
    if (t) {
      const float tv = t[c];
      if (tv > tmax) { tmax = tv; targ = c; }
      dot += x * tv;
      nt += tv * tv;
    }
  }
  smax = warp_max(smax);
  bce = warp_sum(bce);
  float ce = 0.0f, kl = 0.0f;
  int label = 0;
  float lse_s = 0.0f, lse_sT = 0.0f, lse_tT = 0.0f;
  if (t) {
    // first index attaining the maximum (torch.argmax tie rule)
    const float wmax = warp_max(tmax);
    int cand = (tmax == wmax) ? targ : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    // a teacher row that is all NaN / -inf leaves no candidate (no lane compares equal): clamp to class 0 instead of indexing
    // s[INT_MAX] (torch.argmax returns a valid index for such rows too)
    label = (cand < 0 || cand >= p.C) ? 0 : cand;
    tmax = wmax;
    dot = warp_sum(dot); ns = warp_sum(ns); nt = warp_sum(nt);
    float e1 = 0.0f, e2 = 0.0f, e3 = 0.0f;
    const float invT = p.w_kl != 0.0f ? 1.0f / p.T : 0.0f;
    for (int c = lane; c < p.C; c += 32) {
      e1 += __expf(s[c] - smax);
      if (p.w_kl != 0.0f) {
        e2 += __expf((s[c] - smax) * invT);
        e3 += __expf((t[c] - tmax) * invT);
      }
    }
    lse_s = smax + __logf(warp_sum(e1));
    ce = lse_s - s[label];
    if (p.w_kl != 0.0f) {
      lse_sT = smax * invT + __logf(warp_sum(e2));
      lse_tT = tmax * invT + __logf(warp_sum(e3));
      float acc = 0.0f;
      for (int c = lane; c < p.C; c += 32) {
        const float lq = t[c] * invT - lse_tT, lp = s[c] * invT - lse_sT;
        acc += __expf(lq) * (lq - lp);
      }
      kl = warp_sum(acc) * p.T * p.T;
    }
  }
  if (p.dl) {
    const float invT = p.w_kl != 0.0f ? 1.0f / p.T : 0.0f;
    for (int c = lane; c < p.C; c += 32) {
      const float x = s[c];
      float g = p.w_bce * (1.0f / (1.0f + __expf(-x)) - y[c]) * invBC;
      if (t) {
        g += p.w_ce * (__expf(x - lse_s) - (c == label ? 1.0f : 0.0f)) * invB;
        if (p.w_kl != 0.0f) g += p.w_kl * p.T * (__expf(x * invT - lse_sT) - __expf(t[c] * invT - lse_tT)) * invB;
      }
      p.dl[row * p.C + c] = g * p.gscale;
    }
  }
  if (lane == 0) {
    const float lb = bce * invBC, lc = ce * invB, lk = kl * invB;
    atomicAdd(p.losses + 0, p.w_bce * lb + p.w_ce * lc + p.w_kl * lk);
    atomicAdd(p.losses + 1, lb);
    if (t) {
      atomicAdd(p.losses + 2, lc);
      atomicAdd(p.losses + 3, lk);
      if (row == 0) p.losses[4] = dot / (fmaxf(sqrtf(ns), 1e-8f) * fmaxf(sqrtf(nt), 1e-8f));
    }
  }
}

struct HeadParams { const float* z; const float* y; float* prob; float* loss; float* dz; long long G, B; int C; float gscale; };

__global__ void __launch_bounds__(256) pyramid_head_kernel(const HeadParams p) {
  const long long n = p.B * p.C;
  float local = 0.0f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.0f;
    for (long long g = 0; g < p.G; ++g) acc += 1.0f / (1.0f + __expf(-p.z[g * n + i]));
    const float pr = acc / static_cast<float>(p.G);
    p.prob[i] = pr;
    if (p.y) {
      const float yv = p.y[i];
      local += -(yv * fmaxf(__logf(pr), -100.0f) + (1.0f - yv) * fmaxf(__logf(1.0f - pr), -100.0f));
      if (p.dz) {
        const float dldp = (pr - yv) / fmaxf(pr * (1.0f - pr), 1e-12f) / static_cast<float>(n);
        for (long long g = 0; g < p.G; ++g) {
          const float sg = 1.0f / (1.0f + __expf(-p.z[g * n + i]));
          p.dz[g * n + i] = dldp * sg * (1.0f - sg) / static_cast<float>(p.G) * p.gscale;
        }
      }
    }
  }
  if (p.y && p.loss) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.0f) atomicAdd(p.loss, local / static_cast<float>(n));
  }
}

}  // namespace loss
}  // namespace tvt

extern "C" int tvt_distill_loss(const tvt_distill_loss_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->student && a->target && a->losses, "tvt_distill_loss: null pointer");
  TVT_REQUIRE(a->batch > 0 && a->classes > 0 && a->classes <= 65536, "tvt_distill_loss: bad shape");
  TVT_REQUIRE(a->teacher || (a->w_ce == 0.0f && a->w_kl == 0.0f), "tvt_distill_loss: w_ce / w_kl need teacher logits");
  TVT_REQUIRE(a->w_kl == 0.0f || a->temperature > 0.0f, "tvt_distill_loss: temperature must be positive when w_kl != 0");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  loss::Params p{a->student, a->teacher, a->target, a->losses, a->dlogits, a->batch, (int)a->classes,
                 a->w_bce, a->w_ce, a->w_kl, a->temperature, a->grad_scale == 0.0f ? 1.0f : a->grad_scale};
  const int grid = static_cast<int>((a->batch + 3) / 4);
  loss::distill_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("tvt_distill_loss");
}

extern "C" int tvt_pyramid_head(const tvt_pyramid_head_args* a, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(a != nullptr && a->z && a->prob, "tvt_pyramid_head: null pointer");
  TVT_REQUIRE(a->scales > 0 && a->batch > 0 && a->classes > 0, "tvt_pyramid_head: bad shape");
  TVT_REQUIRE(!a->dz || a->target, "tvt_pyramid_head: dz needs target");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  loss::HeadParams p{a->z, a->target, a->prob, a->loss, a->dz, a->scales, a->batch, (int)a->classes, a->grad_scale == 0.0f ? 1.0f : a->grad_scale};
  const long long n = a->batch * a->classes;
  const int grid = static_cast<int>((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  loss::pyramid_head_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("tvt_pyramid_head");
}
