/* A host program in plain C99 against the C-ABI of libmarl_b200 -- no Python, no torch: device memory comes from the CUDA runtime,
 * every call takes raw pointers and sizes.  It builds an HBM replay, inserts synthetic SMAC-3m-shaped episodes from host arrays,
 * creates the recurrent QMIX learner and runs the reference's per-update sequence
 *     sample(B) -> train_policy_on_batch -> soft_target_updates            (offpolicy/runner/rnn/base_runner.py:259-284)
 * printing loss / grad_norm / Q_tot after every step.
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/c_host.c -Loff-policy_b200/lib -lmarl_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/off-policy_b200/lib -o examples/c_host
 *   ./examples/c_host [steps]
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "marl_b200.h"

#define CK(call) do { if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, mx_last_error()); return 1; } } while (0)
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

static float frand(void) { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 5;
  const int N = 3, O = 30, A = 9, S = 48, T = 60, B = 32, E = 128;
  if (!mx_is_cuda_build()) { fprintf(stderr, "not a CUDA build\n"); return 1; }
  cudaStream_t stream;
  CU(cudaStreamCreate(&stream));

  /* ---- replay: one device blob, layout reported by the library ---- */
  mx_replay_cfg rc;
  memset(&rc, 0, sizeof(rc));
  rc.capacity = E; rc.episode_len = T; rc.n_agents = N; rc.obs_dim = O; rc.share_dim = S; rc.act_dim = A;
  rc.use_avail = 1; rc.max_batch = B; rc.per_alpha = 0.6;
  mx_replay_layout lay;
  CK(mx_replay_layout_query(&rc, &lay));
  void* blob;
  CU(cudaMalloc(&blob, (size_t)lay.total_bytes));
  CU(cudaMemsetAsync(blob, 0, (size_t)lay.total_bytes, stream));
  mx_replay* rep;
  CK(mx_replay_create(&rc, blob, stream, &rep));

  /* ---- synthetic episodes in the runner's time-major host layout (rec_buffer.py:146-190), 16 per insert ---- */
  {
    const int n = 16;
    float* obs = malloc(sizeof(float) * (T + 1) * n * N * O);
    float* sh = malloc(sizeof(float) * (T + 1) * n * S);
    float* acts = calloc((size_t)T * n * N * A, sizeof(float));
    float* rew = malloc(sizeof(float) * T * n * N);
    float* dones = calloc((size_t)T * n * N, sizeof(float));
    float* de = calloc((size_t)T * n, sizeof(float));
    float* av = malloc(sizeof(float) * (T + 1) * n * N * A);
    for (int c = 0; c < E; c += n) {
      for (size_t i = 0; i < (size_t)(T + 1) * n * N * O; ++i) obs[i] = frand();
      for (size_t i = 0; i < (size_t)(T + 1) * n * S; ++i) sh[i] = frand();
      for (size_t i = 0; i < (size_t)(T + 1) * n * N * A; ++i) av[i] = 1.f;
      memset(acts, 0, sizeof(float) * T * n * N * A);
      for (size_t i = 0; i < (size_t)T * n * N; ++i) acts[i * A + rand() % A] = 1.f;
      for (int t = 0; t < T; ++t)
        for (int e = 0; e < n; ++e) {
          const float r = frand();
          for (int a = 0; a < N; ++a) rew[((size_t)t * n + e) * N + a] = r;      /* shared team reward */
        }
      mx_episodes ep = {obs, sh, acts, rew, dones, de, av};
      int32_t first;
      CK(mx_replay_insert_async(rep, &ep, n, &first, stream));
      CU(cudaStreamSynchronize(stream));      /* the host arrays are reused for the next insert */
    }
    free(obs); free(sh); free(acts); free(rew); free(dones); free(de); free(av);
  }
  printf("replay: %d episodes, %.1f MB\n", mx_replay_len(rep), lay.total_bytes / 1e6);

  /* ---- learner: flat parameter vectors with the reference's state_dict names ---- */
  mx_qmix_cfg qc;
  memset(&qc, 0, sizeof(qc));
  qc.n_agents = N; qc.obs_dim = O; qc.act_dim = A; qc.state_dim = S; qc.hidden = 64; qc.mixer_hidden = 32; qc.hyper_hidden = 64;
  qc.hyper_layers = 2; qc.episode_len = T; qc.max_batch = B; qc.double_q = 1; qc.use_avail = 1; qc.world_size = 1;
  qc.gamma = 0.99f; qc.huber_delta = 10.f; qc.per_nu = 0.9f; qc.per_eps = 1e-6f;
  qc.lr = 5e-4f; qc.adam_beta1 = 0.9f; qc.adam_beta2 = 0.999f; qc.adam_eps = 1e-5f; qc.max_grad_norm = 10.f; qc.tau = 0.005f;
  mx_param_entry ent[128];
  int64_t P = 0;
  const int n_ent = mx_qmix_param_layout(&qc, ent, 128, &P);
  if (n_ent <= 0) { fprintf(stderr, "param layout: %s\n", mx_last_error()); return 1; }
  float* h_theta = calloc((size_t)P, sizeof(float));
  for (int i = 0; i < n_ent; ++i) {
    const int64_t cnt = (int64_t)ent[i].rows * (ent[i].cols ? ent[i].cols : 1);
    for (int64_t k = 0; k < cnt; ++k) {
      float v;
      if (ent[i].cols == 0) v = strstr(ent[i].name, "weight") ? 1.f : 0.f;      /* 1-D tensors: LayerNorm gains 1, every bias 0 */
      else v = frand() * (float)(1.0 / sqrt((double)ent[i].cols));
      h_theta[ent[i].offset + k] = v;
    }
  }
  float *theta, *theta_tgt, *adam_m, *adam_v;
  CU(cudaMalloc((void**)&theta, P * 4)); CU(cudaMalloc((void**)&theta_tgt, P * 4));
  CU(cudaMalloc((void**)&adam_m, P * 4)); CU(cudaMalloc((void**)&adam_v, P * 4));
  CU(cudaMemcpyAsync(theta, h_theta, P * 4, cudaMemcpyHostToDevice, stream));
  CU(cudaMemcpyAsync(theta_tgt, h_theta, P * 4, cudaMemcpyHostToDevice, stream));
  CU(cudaMemsetAsync(adam_m, 0, P * 4, stream)); CU(cudaMemsetAsync(adam_v, 0, P * 4, stream));
  const int64_t ws_bytes = mx_qmix_workspace_bytes(&qc);
  void* ws;
  CU(cudaMalloc(&ws, (size_t)ws_bytes));
  CU(cudaMemsetAsync(ws, 0, (size_t)ws_bytes, stream));
  mx_qmix* q;
  CK(mx_qmix_create(&qc, theta, theta_tgt, adam_m, adam_v, ws, ws_bytes, &q));
  printf("learner: %d tensors, %lld parameters, %.1f MB workspace\n", n_ent, (long long)P, ws_bytes / 1e6);

  /* ---- the update loop ---- */
  CK(mx_replay_seed(rep, 1u, stream));
  int bad = 0;
  for (int s = 0; s < steps; ++s) {
    mx_batch batch;
    CK(mx_replay_sample_uniform(rep, B, stream));        /* np.random.choice semantics, indices drawn on the device */
    CK(mx_replay_batch(rep, B, &batch));
    CK(mx_qmix_step(q, &batch, stream));                 /* forward, TD target, loss, BPTT, clip, Adam */
    CK(mx_qmix_soft_update(q, stream));
    float info[4];
    CU(cudaMemcpyAsync(info, mx_qmix_info(q), sizeof(info), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    printf("step %d  loss %.6f  grad_norm %.6f  Q_tot %.6f\n", s, info[0], info[1], info[2]);
    if (!(info[0] == info[0]) || !(info[1] == info[1]) || info[0] < 0.f) bad = 1;
  }
  mx_qmix_destroy(q);
  mx_replay_destroy(rep);
  cudaFree(ws); cudaFree(theta); cudaFree(theta_tgt); cudaFree(adam_m); cudaFree(adam_v); cudaFree(blob);
  free(h_theta);
  printf("%s, %lld kernel launches\n", bad ? "FAILED" : "ok", (long long)mx_launch_count());
  return bad;
}
