// Internal host-side structures shared by the translation units of libmarl_b200.
#pragma once
#include "../../include/marl_b200.h"
#include "mx_common.cuh"

#include <functional>
#include <string>

void mx_set_error(const char* fmt, ...);
extern long long g_mx_launches;
extern int g_mx_prof_on;
void mx_prof_mark(const char* name, cudaStream_t s);
extern int g_mx_mixer_split, g_mx_mixer_split_rm, g_mx_overlap, g_mx_overlap_rows, g_mx_mid_fused, g_mx_gru_fwd_rpc, g_mx_gru_bwd_rpc, g_mx_optim_fused, g_mx_gru_threads, g_mx_front_bwd_mma, g_mx_hyper_late, g_mx_side_prio, g_mx_gru_wgrad_split, g_mx_p2p_ll, g_mx_p2p_timeout_ms, g_mx_gru_rows;
int mx_set_option_common(const char* name, int value);   // 0 when `name` was one of the build-independent options
#define MX_COUNT() (++g_mx_launches)
#define MX_MARK(name, s) do { if (g_mx_prof_on) mx_prof_mark((name), (s)); } while (0)

#if MX_EMU
static inline int mx_num_sms() { return 4; }
#define MX_CHECK_LAUNCH(what) 0
#else
int mx_num_sms();
int mx_check_launch(const char* what);
#define MX_CHECK_LAUNCH(what) mx_check_launch(what)
#endif

// whole-step CUDA graph (qmix.cu): the recorded launch sequence, kept for the emulated build which simply re-runs it
struct mx_graph {
#if !MX_EMU
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
#endif
  std::function<int(void*)> seq;
  std::function<void()> after_launch;     // host-side bookkeeping a replay must repeat (e.g. the learner's update counter)
  int n_kernels = 0;
};
int mx_graph_capture_seq(std::function<int(void*)> seq, std::function<void()> after_launch, void* stream, mx_graph** out);

// ---- replay ------------------------------------------------------------------------------------
struct MxReplayState {   // device-resident scalars (off_state)
  int32_t filled;
  int32_t cursor;
  int32_t rng_pos;
  int32_t pad;
  double max_priority;
  double reward_mean, reward_std;
  double per_beta;          // importance-sampling exponent read by the captured PER draw (mx_replay_set_beta); eager calls pass it by value
};

struct mx_replay {
  mx_replay_cfg cfg;
  mx_replay_layout L;
  char* blob;
  int32_t filled, cursor;   // host mirror (inserts are host-driven)
  void* tma = nullptr;      // tensor maps of the fields (gather_tma.cu), null: vectorised gather
};
void* mx_gather_tma_create(mx_replay* r);
void mx_gather_tma_destroy(void* p);
int mx_launch_gather_tma(void* p, const int64_t* idx_dev, int B, cudaStream_t s);     // -1: not available
extern int g_mx_gather_tma;

// ---- agent net / mixer parameter layouts (offsets in floats inside the flat vector) --------------
struct MxNetLayout {       // RNNBase (LN -> fc1 -> LN -> fc2 -> LN -> GRU -> LN) + Linear head
  int in_dim, out_dim;
  int fn_g, fn_b;
  int w1, b1, ln1_g, ln1_b;
  int wh, bh, lnh_g, lnh_b;   // fc_h: registered, unused in forward (mlp.py:21-29); polyak-averaged only
  int w2, b2, ln2_g, ln2_b;
  int wih, whh, bih, bhh;
  int lno_g, lno_b;
  int wq, bq;
  int size;                   // floats, multiple of 4
};
struct MxMixLayout {
  int S, N, ME, HY, layers;
  int w1a, b1a, w1b, b1b;     // hyper_w1: (layers==2) Linear(S,HY) -> ReLU -> Linear(HY,N*ME); (layers==1) only "b" = Linear(S,N*ME)
  int w2a, b2a, w2b, b2b;     // hyper_w2
  int wb1, bb1;               // hyper_b1 Linear(S,ME)
  int wb2a, bb2a, wb2b, bb2b; // hyper_b2 Linear(S,HY) -> ReLU -> Linear(HY,1)
  int size;
};
int mx_net_layout(int in_dim, int out_dim, int base, MxNetLayout* L);
int mx_mix_layout(int S, int N, int ME, int HY, int layers, int base, MxMixLayout* L);

// ---- QMIX learner workspace -----------------------------------------------------------------------
struct MxQmixWs {           // offsets in floats into the workspace
  int64_t gi[2], hall[2];   // [net][M][3H], [net][M][H]
  int64_t u1, u2, st0, st1, st2, sto, gates, hn;   // live-net activations kept for backward
  int64_t qall[2];          // [net][M][A]  (debug + greedy)
  int64_t greedy;           // int32 [M]
  int64_t q_taken, q_next;  // [B*T][N]
  int64_t qtot, qtot_next, err, huberp;  // [B*T]
  int64_t dq_taken;         // [B*T][N]
  int64_t dh_out;           // [M][H]
  int64_t dgi;              // [M][3H]
  int64_t gpart;            // [npart][P]
  int64_t lnpart;           // [2 npart][512]: LayerNorm gain / bias sums of k_front_bwd_tc CTAs (streamed mode, two CTAs per SM)
  int64_t grad;             // [P + 8]   flat gradient numerators + scalars (all-reduce payload)
  int64_t info;             // [8]
  int64_t prio;             // [max_batch]
  int64_t spart;            // [npart][8] per-CTA scalar partials (denominator, loss numerator, sum Q_tot)
  int64_t adam_t;           // double[4]: step count, beta1^t, beta2^t
  int64_t normpart;         // [ceil(P/256)] per-block sums of squares of the reduced gradient numerators
  int64_t xstat;            // float[8]: exchange breakdown accumulated by k_optim_fused (ns: push, wait, sum; launches; max wait)
  int64_t sync;             // uint32[8]: grid-barrier / exchange words of k_optim_fused (zero-initialised with the workspace)
  int64_t tcimg[2];         // pre-split TF32 weight images of the agent front layers (live, target)
  int64_t xin;              // prev_act_inp: packed network input rows [M][round_up(O + A, 4)]
  int64_t tcimgT;           // transposed TF32 weight images for k_front_bwd_tc
  int64_t da2, da1;         // [M][H] each: row gradients handed from k_front_bwd to k_wgrad_tc (option wgrad_tc)
  // split mixer pipeline: per-element hypernet outputs (live: kept for backward; target: forward only) and core's gradients
  int64_t hyp_h1, hyp_h2, hyp_hb;               // live [E][gH]
  int64_t hyp_p1[2], hyp_b1[2], hyp_p2[2], hyp_b2[2];
  int64_t d_q, d_hp, d_p2, d_p1;
  int64_t total;
};

struct mx_qmix {
  mx_qmix_cfg cfg;
  MxNetLayout agent;
  MxMixLayout mix;
  int64_t P;               // padded parameter count
  int npart;               // number of per-CTA gradient partials
  int debug;               // also materialise q_all / greedy (parity tests)
  float *theta, *theta_tgt, *adam_m, *adam_v;
  float* ws;
  int64_t ws_bytes;
  MxQmixWs W;
  int split_ok;            // the split mixer pipeline supports this configuration
  // data-parallel exchange over peer memory (p2p.cu): symmetric blocks of every rank, set by mx_qmix_set_peers
  int p2p_rank = 0, p2p_world = 0;
  float* p2p_blocks[16] = {nullptr};
  uint32_t* p2p_counter = nullptr;
#if !MX_EMU
  // forked branch for the state-only kernels (weight-image prep, mixer hypernets): one non-blocking stream + fork/join events;
  // inside a stream capture the same record/wait calls turn into parallel graph branches
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_prep = nullptr, ev_batch = nullptr, ev_hyper = nullptr, ev_core = nullptr, ev_hbwd = nullptr, ev_gbwd = nullptr;
#endif
  int prep_pending = 0;    // mx_qmix_prefork() already launched the weight-image prep for the coming step
  int imgT_fresh = 0;      // the transposed images (tensor-core backward) were rebuilt with the forward ones for this step
};
int mx_replay_sample_per_state_beta(mx_replay* r, int32_t B, void* stream);   // PER draw whose exponent is the device scalar (captured sequences)
int mx_qmix_prefork(mx_qmix* q, int B, void* stream);   // optional: start the parameter-only work of the next step before its batch is sampled
