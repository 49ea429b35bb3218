// env_handle.cuh — the opaque `tmla_env` handle shared by the translation units that take one (env_kernels.cu, rollout.cu).
#pragma once
#include "common.cuh"

struct tmla_env {
    int task;
    int64_t n;
    uint64_t seed, env_id_base, step_count;
    int device;
    void *buf[4];             // packed SoA planes (device)
    int *err_flag;            // device: set when a kernel saw an out-of-range action
    // staging for the *_host entry points
    void *d_stage, *h_stage;  // device / pinned host, same layout
    size_t stage_bytes;
    cudaStream_t own_stream;
    int32_t *d_ndone;         // device counter of finished episodes in the last step
    int64_t rec_hint;         // records fetched with the first D2H of a host step (1.5x the last count + 256)
    // ordering between the device path (caller's stream) and the host path (own_stream): the last stream a device-path
    // call launched on, and whether anything was launched there since the host path last waited for it
    cudaStream_t dev_stream;
    bool dev_dirty;
    cudaEvent_t dev_evt;
    int act_u8;               // the pinned action stage currently holds uint8 actions (tmla_stage_actions)
    // chunked host step (tmla_step_block_begin / _end): pre-step copy of the state planes (allocated on first use) for the
    // roll-back of a batch whose later chunk is rejected; sequence word of the step in flight (0 = none, -1 = staged only)
    void *shadow[4];
    int pend_seq;
    // optional per-episode log of the policy-driven device path (Monitor rows): {ep_return, ep_length} records
    float2 *ep_log;
    int32_t ep_log_cap;
    int32_t *ep_log_count;
    // captured CUDA graph of the fused policy rollout (rollout.cu), released through the hook by tmla_destroy
    void *rollout_plan;
    void (*rollout_plan_free)(void *);
};

// RAII: run an entry point on the handle's device and give the calling thread its previous device back
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
static inline void mark_device_path(tmla_env *h, cudaStream_t st) { h->dev_stream = st; h->dev_dirty = true; }
