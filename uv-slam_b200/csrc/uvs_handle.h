// uvs_handle.h — the opaque UvsHandle behind the C ABI (internal to libuvs_b200).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/uvs.h"
#include "uvs_device.cuh"
#include "uvs_kernels.h"

namespace uvs {

constexpr size_t ALIGN = 256;
inline size_t align_up(size_t v) { return (v + ALIGN - 1) / ALIGN * ALIGN; }

// bump allocator over one growable device arena (and a mirrored pinned staging buffer)
struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  bool pinned_host = false;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    const size_t want = bytes + bytes / 4 + (1 << 20);
    cudaError_t e = pinned_host ? cudaMallocHost((void **)&base, want) : cudaMalloc((void **)&base, want);
    if (e != cudaSuccess) { base = nullptr; cap = 0; return e; }
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (base) { if (pinned_host) cudaFreeHost(base); else cudaFree(base); }
    base = nullptr; cap = 0;
  }
};

struct Layout {   // byte offsets of one section list; computed twice (input region, work region)
  size_t total = 0;
  // every section STARTS aligned (256 bytes by default; the small per-window tables pack at 16 so that the host-provided
  // part of a single window's upload stays a few KB - uvs_window_upload); the end is not padded
  size_t take(size_t bytes, size_t a = ALIGN) { total = (total + a - 1) / a * a; const size_t o = total; total += std::max<size_t>(bytes, 8); return o; }
};

// byte offsets of the caller-array sections inside the input region (device arena and pinned staging alike)
struct InputOffsets {
  size_t state_end, pfi, pfj, ppt, ppi, ppj, lf, li, lsp, lep, vf, vl, vd, ric, imu_f, idp, idq, idv, idt, iba, ibg, ijac, icov, prJ, prr, prx, total;
};
// Device-resident window (uvs_window.cu): the factor / IMU / prior sections of the input region are written on the device
// from the resident track store instead of being copied from the caller; `fill` is called with the upload's stream
// after the host-provided sections have been enqueued.
struct ResidentHook {
  void *user;
  int line_run_max;   // most observations of one eligible line
  int (*fill)(void *user, UvsHandle *h, char *dev_base, char *stage_base, const InputOffsets &o);
};

}  // namespace uvs

struct UvsHandle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
  uvs::Fork fork{};                           // auxiliary streams + fork / join events (uvs_kernels.h)
  bool concurrent = false;                    // this batch runs independent kernels of a stage side by side
  bool graph_replay = false;                  // uvs_set_graph_replay
  cudaGraphExec_t iter_exec = nullptr;        // one LM iteration of the uploaded batch as a CUDA graph (uvs_solve)
  int64_t iter_launches = 0;                  // kernel launches inside that graph
  std::string err;
  uvs::Arena dev, stage, scratch, hscratch;
  uvs::Dev D{};
  uvs::Params P{};
  UvsOptions opts{};
  bool have_window = false;
  int B = 0, max_d = 0, max_prior_n = 0, packed_limit = 0;
  size_t input_bytes = 0;
  // host copies of the per-window tables
  std::vector<int> frame_off, point_off, line_off, proj_off, lobs_off, vobs_off, imu_off, cam_off, prior_off, pblk_off;
  std::vector<long long> S_off, priorJ_off;
  std::vector<int> win_flags;
  std::vector<int> relo;                      // per window: 1 = the last internal frame is the relocalisation pose (uvs.h, n_relo > 0)
  // offsets of the state sections inside the device arena (buffer 0) for download
  size_t o_pose0 = 0, o_state_bytes = 0;
  size_t in_sb = 0, in_ex = 0, in_td = 0, in_inv = 0, in_ortho = 0, in_ric = 0, in_tic = 0;   // input-region offsets (uvs_upload_state)
  size_t o_reduce = 0, reduce_doubles = 0;   // [Smat | gS | gfull | colsq] block for the multi-GPU exchange
  int64_t launches = 0;
  float last_solve_ms = 0.f, last_sweep_ms = 0.f;
  int n_sweeps = 0;
  int rank = 0, nranks = 1;
  void *nccl_comm = nullptr;                  // ncclComm_t of uvs_comm_init_nccl (uvs_comm.cpp); takes precedence over the callback
  int64_t collectives = 0;                    // all-reduces issued since creation
  UvsAllReduceFn reduce = nullptr;
  void *reduce_user = nullptr;
  int *d_active = nullptr;
  int *h_active = nullptr;
  size_t o_pristine = 0;                      // pristine copy of the uploaded state (uvs_reset_state)
  size_t o_cur = 0, cur_bytes = 0;
  int profiling = 0;
  int max_frames = 0; bool any_ex = false;
  bool use_build3 = false;                    // atomics-free landmark path (uvs_build3.cu)
  bool fused = false;                         // factors evaluated inside the landmark elimination (uvs_lin.cu): no point / line records
  int max_lines = 0;                          // most lines in one window (grid of k_lin_lines)
  // bumped by everything that changes the device state or overwrites the factor records (upload, solve, state upload /
  // reset, the evaluation entry points); uvs_marginalize re-evaluates the factors only when its records are older
  int64_t records_epoch = 0, marg_epoch = -1;
  bool chain_ok = false;                      // speed-bias blocks form a chain in every window: k_chol_chain
  uvs::Build3Layout b3{};
  size_t o_b3 = 0;
  std::vector<cudaEvent_t> stage_ev;          // (UVS_N_STAGES + 1) events per LM iteration
  float stage_ms[UVS_N_STAGES] = {0};
  int stage_iters = 0;
  size_t smem_optin = 0;                      // cudaDeviceProp::sharedMemPerBlockOptin (queried once)
  std::vector<UvsHandle *> children;          // sub-batch handles of uvs_batch_solve_pipelined (own stream + arenas)
  void *resident = nullptr;                   // uvs::ResidentWindow of uvs_window_create (uvs_window.cu)
  const double *last_marg_J = nullptr, *last_marg_r = nullptr;   // device result of the last uvs_marginalize (in `scratch`), n x n and n
  int last_marg_n = 0;
  int64_t h2d_bytes = 0;                      // host-to-device bytes copied through this handle's upload paths since creation
};


namespace uvs {
int handle_fail(UvsHandle *h, int status, const std::string &msg);
int handle_ensure_scratch(UvsHandle *h, size_t bytes);
int handle_ensure_hscratch(UvsHandle *h, size_t bytes);
// uvs_api.cu: upload of one window whose factor sections come from the device-resident store (hook != nullptr)
int handle_upload(UvsHandle *h, int32_t B, const UvsWindow *w, const UvsOptions *opts, const ResidentHook *hook);
// uvs_window.cu
void resident_destroy(UvsHandle *h);
// uvs_comm.cpp: NCCL bound at run time (dlopen of libnccl.so.2; the library has no link-time dependency on it)
int nccl_unique_id(unsigned char id[128]);
int nccl_init_rank(void **comm, const unsigned char id[128], int rank, int nranks);
int nccl_all_reduce_sum(void *comm, double *buf, size_t count, cudaStream_t st);
void nccl_destroy(void *comm);
const char *nccl_error_string(int rc);
}  // namespace uvs
