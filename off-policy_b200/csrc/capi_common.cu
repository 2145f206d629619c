// Error reporting, launch accounting and device queries shared by every entry point.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "mx_internal.h"

static thread_local char g_err[512] = "";
long long g_mx_launches = 0;
#if !MX_EMU
int g_mx_pdl = -1;    // programmatic dependent launch: -1 (default) = for latency-bound learner steps only (rows <= g_mx_pdl_rows; set per step by the
                      // learner), 1 = every launch, 0 = never.  B200, visit 16: 3m 172.0 -> 164.5 us with PDL; 2s3z 463 -> 478, 8m 1126 -> 1156 (the
                      // dependents' prologues then take SM resources from kernels that are throughput-bound)
int g_mx_pdl_rows = 12288;
int g_mx_pdl_auto = 0;      // the learner's per-step decision in automatic mode
int g_mx_pdl_skip_next = 0;
#endif

// runtime options shared by the product and the emulated build (mx_set_option)
int g_mx_p2p_timeout_ms = 10000;      // how long a rank waits for a peer's gradient before it sets the sticky abort word (tests shorten it)
int g_mx_gru_rows = 1;        // rows per CTA of the 128-thread recurrences: 1 (default), 2, 0 = 2 when there are more row-CTAs than two per SM hold at once.
                              // B200, visit 20: two rows per CTA do not pay -- 8m 1 060 vs 1 049 us (k_gru_bwd2 105 -> 86 us alone, k_gru_fwd2 unchanged at 156), 2s3z 438 vs 430
int g_mx_p2p_ll = 1;          // data-parallel exchange inside k_optim_fused: 1 = flag-in-data lines (no fence / counter / flag hop), 0 = slots + per-rank flags
int g_mx_mixer_split = 1;      // 1: split mixer (hypernet-forward / core / hypernet-backward kernels) whenever the forked branch is
                               //    in use; 2: always; 0: always the single fused k_mixer
int g_mx_mixer_split_rm = 0;   // rows per thread of the split mixer's tiles (0 = automatic)
int g_mx_overlap = 1;          // state-only kernels (weight-image prep, mixer hypernets) on a forked branch beside the agent-net
                               // kernels: 1 = when the step is latency-bound (rows <= g_mx_overlap_rows), 2 = always, 0 = never
int g_mx_gru_threads = 0;
int g_mx_side_prio = 0;         // priority of a learner's forked branch (read when the learner is created): 0 default, 1 lower, -1 higher
int g_mx_hyper_late = 0;        // 0 (default): the hypernet branch forks before the front kernel; 1: after it, beside the recurrence -- measured slower
                                // (3m 198.7 vs 184.3 us, MPE 159.7 vs 128.2 us, profiles/r02_option_sweeps.md: the recurrence is the kernel that suffers most from co-residents)
int g_mx_mid_fused = 1;        // 1: k_qhead + k_mix_core + k_qhead_bwd as ONE kernel (k_mid) when the split mixer is in use and no debug
                               //    outputs are requested; 0: three launches
int g_mx_gru_fwd_rpc = 0;      // tuning overrides: sequence rows per CTA of the recurrence kernels (0 = automatic; 1, 2 or 4)
int g_mx_gru_bwd_rpc = 0;
int g_mx_overlap_rows = 1 << 20; // measured on B200 with the fused k_mid in place (profiles/r02_option_sweeps.md, visit 12): forked branch 3m +16 %, 2s3z (19 360 rows) +12 %,
                               // 8m (61 952 rows) +7 %; round 1 (without k_mid for 8 agents) had 8m at -8 % and a 12 288-row limit.  Beyond 1M rows: unmeasured, serial
int mx_set_option_common(const char* name, int value) {
  if (!strcmp(name, "mixer_split")) { g_mx_mixer_split = value; return 0; }
  if (!strcmp(name, "mixer_split_rm")) { g_mx_mixer_split_rm = value; return 0; }
  if (!strcmp(name, "overlap")) { g_mx_overlap = value; return 0; }
  if (!strcmp(name, "overlap_rows")) { g_mx_overlap_rows = value; return 0; }
  if (!strcmp(name, "mid_fused")) { g_mx_mid_fused = value; return 0; }
  if (!strcmp(name, "optim_fused")) { g_mx_optim_fused = value; return 0; }
  if (!strcmp(name, "side_prio")) { g_mx_side_prio = value; return 0; }
  if (!strcmp(name, "hyper_late")) { g_mx_hyper_late = value; return 0; }
  if (!strcmp(name, "front_bwd_mma")) { g_mx_front_bwd_mma = value; return 0; }
  if (!strcmp(name, "gru_wgrad_split")) { g_mx_gru_wgrad_split = value; return 0; }
  if (!strcmp(name, "p2p_ll")) { g_mx_p2p_ll = value; return 0; }
  if (!strcmp(name, "gru_rows")) { g_mx_gru_rows = value; return 0; }
  if (!strcmp(name, "p2p_timeout_ms")) { g_mx_p2p_timeout_ms = value; return 0; }
#if !MX_EMU
  if (!strcmp(name, "smem_carveout")) { g_mx_smem_carveout = value; return 0; }
#else
  if (!strcmp(name, "smem_carveout")) return 0;
#endif
  if (!strcmp(name, "gather_tma")) { g_mx_gather_tma = value; return 0; }      // 1: episode gather on the TMA unit (default), 0: vectorised loads
  if (!strcmp(name, "gru_threads")) { g_mx_gru_threads = value; return 0; }      // 0: by size, 128 / 256: force the recurrence kernels' CTA width
  if (!strcmp(name, "gru_fwd_rpc")) { g_mx_gru_fwd_rpc = value; return 0; }
  if (!strcmp(name, "gru_bwd_rpc")) { g_mx_gru_bwd_rpc = value; return 0; }
  return -1;
}

void mx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mx_last_error(void) { return g_err; }
extern "C" int mx_abi_version(void) { return MX_ABI_VERSION; }

// ---- host fences: "has the stream consumed this pinned staging buffer yet?" -------------------------------------------------------------
// The drop-in buffers reuse a few pinned host blocks (insert staging, sampled-index ring); a block may be rewritten only after the copy
// that read it has completed.  A pool of timing-less CUDA events behind three tiny calls: a torch.cuda.Event().record() costs the
// host ~8 us per call (it resolves the current stream object first), these ~1 us.
#define MX_MAX_FENCES 256
#if !MX_EMU
static cudaEvent_t g_fence[MX_MAX_FENCES];
static unsigned char g_fence_set[MX_MAX_FENCES];
#endif
static int g_fence_n = 0;
extern "C" int mx_host_fence_alloc(void) {
  if (g_fence_n >= MX_MAX_FENCES) { mx_set_error("mx_host_fence_alloc: out of fences"); return -1; }
#if !MX_EMU
  if (cudaEventCreateWithFlags(&g_fence[g_fence_n], cudaEventDisableTiming) != cudaSuccess) { mx_set_error("mx_host_fence_alloc: cudaEventCreate failed"); return -1; }
  g_fence_set[g_fence_n] = 0;
#endif
  return g_fence_n++;
}
extern "C" int mx_host_fence_record(int id, void* stream) {
  if (id < 0 || id >= g_fence_n) { mx_set_error("mx_host_fence_record: bad fence"); return 1; }
#if !MX_EMU
  if (cudaEventRecord(g_fence[id], (cudaStream_t)stream) != cudaSuccess) { mx_set_error("mx_host_fence_record: cudaEventRecord failed"); return 1; }
  g_fence_set[id] = 1;
#else
  (void)stream;
#endif
  return 0;
}
extern "C" int mx_host_fence_wait(int id) {       // returns once everything enqueued before the last record of this fence has completed
  if (id < 0 || id >= g_fence_n) { mx_set_error("mx_host_fence_wait: bad fence"); return 1; }
#if !MX_EMU
  if (g_fence_set[id]) {
    const cudaError_t e = cudaEventSynchronize(g_fence[id]);      // also the first place an asynchronous fault of earlier work surfaces
    if (e != cudaSuccess) { mx_set_error("mx_host_fence_wait: cudaEventSynchronize failed: %s", cudaGetErrorString(e)); return 1; }
  }
#endif
  return 0;
}
extern "C" int64_t mx_sizeof(const char* n) {
  if (!n) return -1;
#define MX_SZ(T) if (!strcmp(n, #T)) return (int64_t)sizeof(T)
  MX_SZ(mx_batch); MX_SZ(mx_replay_cfg); MX_SZ(mx_replay_layout); MX_SZ(mx_qmix_cfg); MX_SZ(mx_maddpg_cfg); MX_SZ(mx_param_entry);
  MX_SZ(mx_policy_step_args); MX_SZ(mx_episodes);
#undef MX_SZ
  return -1;
}
extern "C" int mx_is_cuda_build(void) { return MX_EMU ? 0 : 1; }
extern "C" int64_t mx_launch_count(void) { return g_mx_launches; }

int g_mx_prof_on = 0;
#include <string>
#include <vector>
#if MX_EMU
// CPU-emulated unit-test build: the same marks, stamped with the host clock (kernels run synchronously there)
#include <chrono>
static std::vector<std::pair<std::string, double>> g_marks;
static double g_prof_start;
static double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
void mx_prof_mark(const char* name, cudaStream_t) { g_marks.emplace_back(name, emu_now_ms()); }
extern "C" int mx_profile_begin(void*) {
  g_marks.clear();
  g_prof_start = emu_now_ms();
  g_mx_prof_on = 1;
  return 0;
}
extern "C" int mx_profile_end(void*, char* names_buf, int32_t buf_len, float* ms, int32_t max_n) {
  g_mx_prof_on = 0;
  std::string names;
  double prev = g_prof_start;
  int n = 0;
  for (auto& m : g_marks) {
    if (n < max_n) {
      ms[n] = (float)(m.second - prev);
      if (n) names += ";";
      names += m.first;
      ++n;
    }
    prev = m.second;
  }
  if (names_buf && buf_len > 0) snprintf(names_buf, buf_len, "%s", names.c_str());
  g_marks.clear();
  return n;
}
#else
static std::vector<std::pair<std::string, cudaEvent_t>> g_marks;
static cudaEvent_t g_prof_start;
void mx_prof_mark(const char* name, cudaStream_t s) {
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  g_marks.emplace_back(name, e);
}
extern "C" int mx_profile_begin(void* stream) {
  for (auto& m : g_marks) cudaEventDestroy(m.second);
  g_marks.clear();
  cudaEventCreate(&g_prof_start);
  cudaEventRecord(g_prof_start, (cudaStream_t)stream);
  g_mx_prof_on = 1;
  return 0;
}
extern "C" int mx_profile_end(void* stream, char* names_buf, int32_t buf_len, float* ms, int32_t max_n) {
  g_mx_prof_on = 0;
  cudaStreamSynchronize((cudaStream_t)stream);
  std::string names;
  cudaEvent_t prev = g_prof_start;
  int n = 0;
  for (auto& m : g_marks) {
    if (n < max_n) {
      float t = 0.f;
      cudaEventElapsedTime(&t, prev, m.second);
      ms[n] = t;
      if (n) names += ";";
      names += m.first;
      ++n;
    }
    prev = m.second;
  }
  if (names_buf && buf_len > 0) snprintf(names_buf, buf_len, "%s", names.c_str());
  for (auto& m : g_marks) cudaEventDestroy(m.second);
  g_marks.clear();
  cudaEventDestroy(g_prof_start);
  return n;
}
#endif

#if !MX_EMU
int g_mx_smem_carveout = 100;
void mx_prefer_carveout(const void* kern) {
  static const void* seen[128];
  static int nseen = 0, applied = -2;
  if (applied != g_mx_smem_carveout) { nseen = 0; applied = g_mx_smem_carveout; }      // option changed: re-apply to every kernel
  for (int i = 0; i < nseen; ++i) if (seen[i] == kern) return;
  if (nseen < 128) seen[nseen++] = kern;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, g_mx_smem_carveout < 0 ? -1 : g_mx_smem_carveout);
}
int mx_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}
int mx_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    mx_set_error("CUDA error after %s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}
#endif
