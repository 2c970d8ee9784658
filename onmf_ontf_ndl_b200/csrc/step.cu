// The fused online step: ONE host call enqueues everything `Online_NTF.step` (reference src/ontf.py:117-154) does for a
// minibatch -- dictionary update with the OLD aggregates, FP64 Gram + TF32 split of the new dictionary (side stream);
// covariances, LARS-lasso codes, surrogate partial sums (main stream); blend into A, B (side stream) -- with the
// cross-stream ordering expressed through CUDA events owned by an opaque plan.  No device memory is allocated and
// nothing synchronises the host.  A multi-GPU caller all-reduces P[cur] on the side stream between
// onmf_step_launch() and onmf_step_finish(); a single-GPU caller uses onmf_step().
//
// Ordering (one-step lag of the reference: update_dict at step t reads A_{t-1}, B_{t-1}, src/ontf.py:151):
//   side : wait(code_{t-1}) -> BCD(W[cur] -> W[cur^1]) -> Gram, split -> record(W_t)
//   main : [split X] -> cov(W[cur]) -> [wait(AB_{t-1}) if hold_coder] -> LARS -> [split H] -> partial sums -> record(P_t, code_t)
//   side : wait(P_t) -> (all-reduce by the caller) -> blend -> record(AB_t)
//   main : wait(W_t)                                       -- the next coding needs the new dictionary only
#include <new>

#include "common.cuh"

// one instantiated CUDA graph of the whole step for a fixed (minibatch pointer, codes pointer, n, cur, buffers)
struct StepGraph {
  const void* Xt;
  const void* codes;
  long long n;
  int cur;
  unsigned long long sig;     // hash of the buffer descriptor the graph was captured with
  cudaGraphExec_t exec;       // nullptr: key seen once, not captured yet
  long long kernels;          // kernels inside the graph (launch accounting)
  unsigned long long last_use;
};
constexpr int STEP_GRAPHS = 8;
constexpr int W_RING = 4096;

struct onmf_step_plan {
  cudaEvent_t ev_P, ev_W, ev_code, ev_AB;
  cudaEvent_t ev_g0, ev_g1, ev_g2, ev_pre;   // fork / partial-sums / join inside a captured step; pre-launch join
  int tslots;                 // LARS timing ring (0 = off)
  long long tcount;
  cudaEvent_t* t0;
  cudaEvent_t* t1;
  long long launches;         // kernels launched through this plan
  StepGraph graphs[STEP_GRAPHS];
  unsigned long long use_clock;
  cudaStream_t cap;           // capture origin (the caller's main stream may be the legacy default stream, which cannot capture)
  double* w_ring;             // pinned host ring of blend weights (copied to buffers->w_dev before each graph launch)
  unsigned long long w_count;
  long long graph_steps;      // steps that ran as a graph replay
};

using namespace onmf;

extern "C" int onmf_step_plan_create(onmf_step_plan** out, int timing_slots) {
  if (!out || timing_slots < 0 || timing_slots > 4096) return fail(ONMF_E_ARG, "step_plan_create: bad argument");
  onmf_step_plan* p = new (std::nothrow) onmf_step_plan();
  if (!p) return fail(ONMF_E_ARG, "step_plan_create: out of host memory");
  p->tslots = timing_slots; p->tcount = 0; p->t0 = p->t1 = nullptr; p->launches = 0;
  memset(p->graphs, 0, sizeof(p->graphs));
  p->use_clock = 0; p->w_ring = nullptr; p->w_count = 0; p->graph_steps = 0;
  ONMF_CUDA(cudaHostAlloc((void**)&p->w_ring, W_RING * sizeof(double), cudaHostAllocDefault));
  ONMF_CUDA(cudaStreamCreateWithFlags(&p->cap, cudaStreamNonBlocking));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_g0, cudaEventDisableTiming));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_g1, cudaEventDisableTiming));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_g2, cudaEventDisableTiming));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_pre, cudaEventDisableTiming));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_P, cudaEventDisableTiming));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_W, cudaEventDisableTiming));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_code, cudaEventDisableTiming));
  ONMF_CUDA(cudaEventCreateWithFlags(&p->ev_AB, cudaEventDisableTiming));
  if (timing_slots > 0) {
    p->t0 = new (std::nothrow) cudaEvent_t[timing_slots];
    p->t1 = new (std::nothrow) cudaEvent_t[timing_slots];
    if (!p->t0 || !p->t1) return fail(ONMF_E_ARG, "step_plan_create: out of host memory");
    for (int i = 0; i < timing_slots; ++i) {
      ONMF_CUDA(cudaEventCreate(&p->t0[i]));
      ONMF_CUDA(cudaEventCreate(&p->t1[i]));
    }
  }
  *out = p;
  return ONMF_OK;
}

extern "C" int onmf_step_plan_destroy(onmf_step_plan* p) {
  if (!p) return ONMF_OK;
  cudaEventDestroy(p->ev_P); cudaEventDestroy(p->ev_W); cudaEventDestroy(p->ev_code); cudaEventDestroy(p->ev_AB);
  cudaEventDestroy(p->ev_g0); cudaEventDestroy(p->ev_g1); cudaEventDestroy(p->ev_g2); cudaEventDestroy(p->ev_pre);
  for (int i = 0; i < STEP_GRAPHS; ++i)
    if (p->graphs[i].exec) cudaGraphExecDestroy(p->graphs[i].exec);
  if (p->w_ring) cudaFreeHost(p->w_ring);
  if (p->cap) cudaStreamDestroy(p->cap);
  for (int i = 0; i < p->tslots; ++i) { cudaEventDestroy(p->t0[i]); cudaEventDestroy(p->t1[i]); }
  delete[] p->t0;
  delete[] p->t1;
  delete p;
  return ONMF_OK;
}

extern "C" int onmf_step_plan_mark_state(onmf_step_plan* p, void* main_stream) {
  if (!p) return fail(ONMF_E_ARG, "step_plan_mark_state: null plan");
  ONMF_CUDA(cudaEventRecord(p->ev_code, (cudaStream_t)main_stream));
  return ONMF_OK;
}

extern "C" long long onmf_step_plan_launches(const onmf_step_plan* p) { return p ? p->launches : 0; }
extern "C" long long onmf_step_plan_graph_steps(const onmf_step_plan* p) { return p ? p->graph_steps : 0; }

extern "C" int onmf_step_plan_reset_timing(onmf_step_plan* p) {
  if (!p) return fail(ONMF_E_ARG, "step_plan_reset_timing: null plan");
  p->tcount = 0;
  return ONMF_OK;
}

extern "C" int onmf_step_plan_lars_ms(onmf_step_plan* p, float* out, int max_out, int* n_out) {
  if (!p || !out || !n_out) return fail(ONMF_E_ARG, "step_plan_lars_ms: bad argument");
  long long have = p->tcount < p->tslots ? p->tcount : p->tslots;
  int n = (int)(have < max_out ? have : max_out);
  const long long first = p->tcount - have;
  for (int i = 0; i < n; ++i) {
    const int s = (int)((first + i) % p->tslots);
    ONMF_CUDA(cudaEventSynchronize(p->t1[s]));
    ONMF_CUDA(cudaEventElapsedTime(&out[i], p->t0[s], p->t1[s]));
  }
  *n_out = n;
  return ONMF_OK;
}

static int lars_launch_count(int k) { return k <= 32 ? 2 : k <= 64 ? 5 : 6; }   // pad + tier chain + hint (lars.cu)

static int check_buffers(const onmf_step_buffers* b, bool need_code_bufs) {
  if (!b) return fail(ONMF_E_ARG, "step: null buffers");
  if (b->dtype != ONMF_F32 && b->dtype != ONMF_F64) return fail(ONMF_E_ARG, "step: bad dtype");
  if (b->d <= 0 || b->k <= 0) return fail(ONMF_E_ARG, "step: bad shape");
  if (!b->W[0] || !b->W[1] || !b->G[0] || !b->G[1] || !b->A || !b->B || !b->P[0] || !b->P[1] || !b->ws_gram)
    return fail(ONMF_E_ARG, "step: null state buffer");
  if (b->use_tc && (b->dtype != ONMF_F32 || !b->Whi[0] || !b->Whi[1] || !b->Wlo[0] || !b->Wlo[1]))
    return fail(ONMF_E_ARG, "step: tensor-core path needs fp32 and the split dictionary buffers");
  if (need_code_bufs && (!b->Ht || !b->ws_sur)) return fail(ONMF_E_ARG, "step: null minibatch buffer");
  if (b->track_C && (!b->C || !b->P2)) return fail(ONMF_E_ARG, "step: track_C needs C and P2");
  return ONMF_OK;
}

extern "C" int onmf_step_launch(onmf_step_plan* p, const onmf_step_buffers* b, const void* Xt, const void* codes,
                                int64_t n, int cur) {
  if (!p) return fail(ONMF_E_ARG, "step_launch: null plan");
  int rc = check_buffers(b, n > 0);
  if (rc) return rc;
  if (n < 0 || (cur != 0 && cur != 1)) return fail(ONMF_E_ARG, "step_launch: bad argument");
  const bool presplit = (Xt == nullptr);
  if (n > 0 && presplit && !b->use_tc) return fail(ONMF_E_ARG, "step_launch: Xt = NULL needs the tensor-core path");
  if (n > 0 && presplit && b->track_C) return fail(ONMF_E_ARG, "step_launch: track_C needs the unsplit minibatch");
  cudaStream_t main = (cudaStream_t)b->main_stream, side = (cudaStream_t)b->side_stream;
  const int dt = b->dtype, d = b->d, k = b->k, nx = cur ^ 1;
  const size_t esz = dt == ONMF_F64 ? 8 : 4;

  // ---- side: dictionary update with the OLD aggregates, then everything the coder derives from the dictionary ----
  ONMF_CUDA(cudaStreamWaitEvent(side, p->ev_code, 0));
  // (the large-dictionary fallback of the update borrows the Gram workspace: same stream, used one after the other)
  if ((rc = onmf_update_dict_ws(dt, b->W[cur], b->A, b->B, d, k, b->W[nx], b->ws_gram, b->ws_gram_bytes, side))) return rc;
  if ((rc = onmf_gram_f64(dt, b->W[nx], d, k, b->G[nx], nullptr, b->ws_gram, b->ws_gram_bytes, side))) return rc;
  p->launches += 3;
  if (b->use_tc) {
    if ((rc = onmf_split_tf32(b->W[nx], b->Whi[nx], b->Wlo[nx], (int64_t)d * k, side))) return rc;
    p->launches += 1;
  }
  ONMF_CUDA(cudaEventRecord(p->ev_W, side));

  // ---- main: code this minibatch with W[cur], partial sums ----
  if (b->track_C) ONMF_CUDA(cudaStreamWaitEvent(main, p->ev_AB, 0));     // P2 is single-buffered (AB is recorded after C's blend)
  if (n > 0) {
    const void* Hcodes = codes ? codes : b->Ht;
    if (b->use_tc && !presplit) {
      if ((rc = onmf_split_tf32(Xt, b->Xhi, b->Xlo, n * d, main))) return rc;
      p->launches += 1;
    }
    if (!codes) {
      if (!b->Ct || !b->ws_lars) return fail(ONMF_E_ARG, "step_launch: null coder buffer");
      if (b->use_tc) rc = onmf_cov_tc(b->Xhi, b->Xlo, n, d, b->Whi[cur], b->Wlo[cur], k, b->Ct, main);
      else rc = onmf_cov(dt, Xt, n, d, b->W[cur], k, b->Ct, main);
      if (rc) return rc;
      // the dictionary update is one thread-block cluster: it is only placed while a group of SMs in one GPC is free,
      // i.e. before the persistent coder has spread over the GPU -- hold the coder until the blend it follows is done
      if (b->hold_coder) ONMF_CUDA(cudaStreamWaitEvent(main, p->ev_AB, 0));
      int rsv = b->reserve_sms >= 0 ? b->reserve_sms : ((long long)n * k <= 131072LL * 256 ? 8 : 0);
      const int saved = g_lars_reserved_sms;
      g_lars_reserved_sms = rsv;
      if (p->tslots > 0) ONMF_CUDA(cudaEventRecord(p->t0[p->tcount % p->tslots], main));
      if (dt == ONMF_F32)
        rc = onmf_lasso_lars_g64(dt, b->G[cur], b->Ct, n, k, d, b->alpha, b->max_iter, b->Ht, b->ws_lars, b->ws_lars_bytes,
                                 b->stats, -1, main);
      else
        rc = onmf_lasso_lars_ex(dt, b->G[cur], b->Ct, n, k, d, b->alpha, b->max_iter, b->Ht, b->ws_lars, b->ws_lars_bytes,
                                b->stats, -1, main);
      g_lars_reserved_sms = saved;
      if (rc) return rc;
      if (p->tslots > 0) {
        ONMF_CUDA(cudaEventRecord(p->t1[p->tcount % p->tslots], main));
        ++p->tcount;
      }
      p->launches += 1 + lars_launch_count(k);
    }
    if (b->use_tc) {
      if ((rc = onmf_split_tf32(Hcodes, b->Hhi, b->Hlo, n * k, main))) return rc;
      if ((rc = onmf_surrogate_partial_tc(b->Hhi, b->Hlo, b->Xhi, b->Xlo, n, k, d, b->P[cur], b->ws_sur, b->ws_sur_bytes, main)))
        return rc;
      p->launches += 5;
    } else {
      if ((rc = onmf_surrogate_partial(dt, Hcodes, Xt, n, k, d, b->P[cur], b->ws_sur, b->ws_sur_bytes, main))) return rc;
      p->launches += 3;
    }
    if (b->track_C) {
      if ((rc = onmf_xxt_partial(dt, Xt, n, d, b->P2, b->ws_sur, b->ws_sur_bytes, main))) return rc;
      p->launches += 2;
    }
  } else {
    ONMF_CUDA(cudaMemsetAsync(b->P[cur], 0, (size_t)k * (k + d) * esz, main));
    if (b->track_C) ONMF_CUDA(cudaMemsetAsync(b->P2, 0, (size_t)d * d * esz, main));
  }
  ONMF_CUDA(cudaEventRecord(p->ev_P, main));
  ONMF_CUDA(cudaEventRecord(p->ev_code, main));
  // ---- side: the blend (and the caller's all-reduce before it) needs this step's partial sums ----
  ONMF_CUDA(cudaStreamWaitEvent(side, p->ev_P, 0));
  return ONMF_OK;
}

extern "C" int onmf_step_finish(onmf_step_plan* p, const onmf_step_buffers* b, double w, int cur) {
  if (!p) return fail(ONMF_E_ARG, "step_finish: null plan");
  int rc = check_buffers(b, false);
  if (rc) return rc;
  cudaStream_t main = (cudaStream_t)b->main_stream, side = (cudaStream_t)b->side_stream;
  if ((rc = onmf_surrogate_blend(b->dtype, b->P[cur], b->k, b->d, w, b->A, b->B, side))) return rc;
  p->launches += 1;
  if (b->track_C) {
    if ((rc = onmf_axpby(b->dtype, (int64_t)b->d * b->d, w, b->P2, 1.0 - w, b->C, side))) return rc;
    p->launches += 1;
  }
  ONMF_CUDA(cudaEventRecord(p->ev_AB, side));
  ONMF_CUDA(cudaStreamWaitEvent(main, p->ev_W, 0));     // the next coding needs the new dictionary only
  return ONMF_OK;
}

extern "C" int onmf_step(onmf_step_plan* p, const onmf_step_buffers* b, const void* Xt, const void* codes, int64_t n,
                         double w, int cur) {
  int rc = onmf_step_launch(p, b, Xt, codes, n, cur);
  if (rc) return rc;
  return onmf_step_finish(p, b, w, cur);
}

// ------------------------------------------------------------------------------------------------------------------
// The same step as ONE CUDA graph (single GPU).  A step is a fork-join: [dictionary update, Gram, split] on the side
// branch next to [covariances, coder, partial sums] on the main branch, then the blend that needs both -- and the next
// step's coding needs nothing but the new dictionary, so consecutive steps are consecutive graph launches on the main
// stream.  For the small configurations (BASELINE configs[0..3]: a step is ~15 dependent launches of a few microseconds
// each) this removes the per-launch CPU cost and the inter-kernel gaps.  The blend weight w = t^-beta changes every
// step: it is read from device memory (buffers->w_dev), refreshed by a stream-ordered 8-byte copy before each launch.
// Graphs are cached per (minibatch pointer, codes pointer, n, cur, buffer descriptor); a key is captured the second
// time it is seen, so one-off calls never pay for capture + instantiation.
// ------------------------------------------------------------------------------------------------------------------
static unsigned long long buffers_sig(const onmf_step_buffers* b) {
  unsigned long long h = 1469598103934665603ULL;
  const unsigned char* q = reinterpret_cast<const unsigned char*>(b);
  for (size_t i = 0; i < sizeof(*b); ++i) { h ^= q[i]; h *= 1099511628211ULL; }
  return h;
}

static int enqueue_captured(onmf_step_plan* p, const onmf_step_buffers* b, const void* Xt, const void* codes, int64_t n, int cur,
                            long long* kernels) {
  // the main branch is captured on the plan's own stream (kernels do not care which stream recorded them; the graph is
  // launched on the caller's main stream afterwards)
  cudaStream_t main = p->cap, side = (cudaStream_t)b->side_stream;
  const int dt = b->dtype, d = b->d, k = b->k, nx = cur ^ 1;
  const size_t esz = dt == ONMF_F64 ? 8 : 4;
  const bool presplit = (Xt == nullptr);
  int rc;
  long long kn = 0;
  ONMF_CUDA(cudaEventRecord(p->ev_g0, main));
  ONMF_CUDA(cudaStreamWaitEvent(side, p->ev_g0, 0));                      // fork
  if ((rc = onmf_update_dict_ws(dt, b->W[cur], b->A, b->B, d, k, b->W[nx], b->ws_gram, b->ws_gram_bytes, side))) return rc;
  if ((rc = onmf_gram_f64(dt, b->W[nx], d, k, b->G[nx], nullptr, b->ws_gram, b->ws_gram_bytes, side))) return rc;
  kn += 3;
  if (b->use_tc) {
    if ((rc = onmf_split_tf32(b->W[nx], b->Whi[nx], b->Wlo[nx], (int64_t)d * k, side))) return rc;
    kn += 1;
  }
  if (n > 0) {
    const void* Hcodes = codes ? codes : b->Ht;
    if (b->use_tc && !presplit) {
      if ((rc = onmf_split_tf32(Xt, b->Xhi, b->Xlo, n * d, main))) return rc;
      kn += 1;
    }
    if (!codes) {
      if (!b->Ct || !b->ws_lars) return fail(ONMF_E_ARG, "step_graph: null coder buffer");
      if (b->use_tc) rc = onmf_cov_tc(b->Xhi, b->Xlo, n, d, b->Whi[cur], b->Wlo[cur], k, b->Ct, main);
      else rc = onmf_cov(dt, Xt, n, d, b->W[cur], k, b->Ct, main);
      if (rc) return rc;
      const int saved = g_lars_reserved_sms;
      g_lars_reserved_sms = b->reserve_sms >= 0 ? b->reserve_sms : 0;
      if (dt == ONMF_F32)
        rc = onmf_lasso_lars_g64(dt, b->G[cur], b->Ct, n, k, d, b->alpha, b->max_iter, b->Ht, b->ws_lars, b->ws_lars_bytes,
                                 b->stats, -1, main);
      else
        rc = onmf_lasso_lars_ex(dt, b->G[cur], b->Ct, n, k, d, b->alpha, b->max_iter, b->Ht, b->ws_lars, b->ws_lars_bytes,
                                b->stats, -1, main);
      g_lars_reserved_sms = saved;
      if (rc) return rc;
      kn += 1 + lars_launch_count(k);
    }
    if (b->use_tc) {
      if ((rc = onmf_split_tf32(Hcodes, b->Hhi, b->Hlo, n * k, main))) return rc;
      if ((rc = onmf_surrogate_partial_tc(b->Hhi, b->Hlo, b->Xhi, b->Xlo, n, k, d, b->P[cur], b->ws_sur, b->ws_sur_bytes, main)))
        return rc;
      kn += 5;
    } else {
      if ((rc = onmf_surrogate_partial(dt, Hcodes, Xt, n, k, d, b->P[cur], b->ws_sur, b->ws_sur_bytes, main))) return rc;
      kn += 3;
    }
  } else {
    ONMF_CUDA(cudaMemsetAsync(b->P[cur], 0, (size_t)k * (k + d) * esz, main));
  }
  ONMF_CUDA(cudaEventRecord(p->ev_g1, main));
  ONMF_CUDA(cudaStreamWaitEvent(side, p->ev_g1, 0));
  if ((rc = onmf_surrogate_blend_dev(dt, b->P[cur], k, d, b->w_dev, b->A, b->B, side))) return rc;
  kn += 1;
  ONMF_CUDA(cudaEventRecord(p->ev_g2, side));
  ONMF_CUDA(cudaStreamWaitEvent(main, p->ev_g2, 0));                      // join
  *kernels = kn;
  return ONMF_OK;
}

extern "C" int onmf_step_graph(onmf_step_plan* p, const onmf_step_buffers* b, const void* Xt, const void* codes, int64_t n,
                               double w, int cur) {
  if (!p) return fail(ONMF_E_ARG, "step_graph: null plan");
  int rc = check_buffers(b, n > 0);
  if (rc) return rc;
  if (n < 0 || (cur != 0 && cur != 1)) return fail(ONMF_E_ARG, "step_graph: bad argument");
  const bool presplit = (Xt == nullptr);
  if (n > 0 && presplit && !b->use_tc) return fail(ONMF_E_ARG, "step_graph: Xt = NULL needs the tensor-core path");
  // not expressible as this graph: the d x d aggregate (host-side w in its blend), coder timing events, multi-GPU hold
  if (!b->w_dev || b->track_C || p->tslots > 0 || b->hold_coder) return onmf_step(p, b, Xt, codes, n, w, cur);
  cudaStream_t main = (cudaStream_t)b->main_stream, side = (cudaStream_t)b->side_stream;
  const unsigned long long sig = buffers_sig(b);
  StepGraph* e = nullptr;
  for (int i = 0; i < STEP_GRAPHS; ++i) {
    StepGraph& g = p->graphs[i];
    if (g.last_use && g.Xt == Xt && g.codes == codes && g.n == n && g.cur == cur && g.sig == sig) { e = &g; break; }
  }
  if (!e) {                                       // first sight: remember the key, run the stream schedule
    StepGraph* v = &p->graphs[0];
    for (int i = 1; i < STEP_GRAPHS; ++i)
      if (p->graphs[i].last_use < v->last_use) v = &p->graphs[i];
    if (v->exec) cudaGraphExecDestroy(v->exec);
    v->Xt = Xt; v->codes = codes; v->n = n; v->cur = cur; v->sig = sig; v->exec = nullptr; v->kernels = 0;
    v->last_use = ++p->use_clock;
    return onmf_step(p, b, Xt, codes, n, w, cur);
  }
  e->last_use = ++p->use_clock;
  if (!e->exec) {                                 // second sight: capture
    // everything the side stream still has in flight from stream-scheduled steps must be ordered before the capture's
    // main-stream origin, and nothing may be captured that waits on events recorded outside the capture
    ONMF_CUDA(cudaEventRecord(p->ev_pre, side));
    ONMF_CUDA(cudaStreamWaitEvent(main, p->ev_pre, 0));
    ONMF_CUDA(cudaStreamBeginCapture(p->cap, cudaStreamCaptureModeThreadLocal));
    long long kn = 0;
    rc = enqueue_captured(p, b, Xt, codes, n, cur, &kn);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(p->cap, &graph);
    if (rc || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc) return rc;
      return cuda_fail(ce, "cudaStreamEndCapture");
    }
    ce = cudaGraphInstantiate(&e->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { e->exec = nullptr; return cuda_fail(ce, "cudaGraphInstantiate"); }
    e->kernels = kn;
  }
  ONMF_CUDA(cudaEventRecord(p->ev_pre, side));
  ONMF_CUDA(cudaStreamWaitEvent(main, p->ev_pre, 0));
  double* slot = &p->w_ring[p->w_count++ % W_RING];
  *slot = w;
  ONMF_CUDA(cudaMemcpyAsync(b->w_dev, slot, sizeof(double), cudaMemcpyHostToDevice, main));
  ONMF_CUDA(cudaGraphLaunch(e->exec, main));
  // leave the stream-schedule events in a consistent state (a later stream-scheduled step, set_state or flush waits on them)
  ONMF_CUDA(cudaEventRecord(p->ev_P, main));
  ONMF_CUDA(cudaEventRecord(p->ev_code, main));
  ONMF_CUDA(cudaEventRecord(p->ev_W, main));
  ONMF_CUDA(cudaEventRecord(p->ev_AB, main));
  p->launches += e->kernels;
  ++p->graph_steps;
  return ONMF_OK;
}
