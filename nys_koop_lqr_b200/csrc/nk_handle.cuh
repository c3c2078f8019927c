// nk_handle.cuh -- the opaque handle behind include/nk_b200.h: device, grow-only workspaces, error text.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "nk_gram.cuh"
#include "../../include/nk_b200.h"

struct nk_devbuf {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct nk_handle {
    int device = 0;
    int sm_count = 0;
    std::string err;
    long long launches = 0;

    // ---- fused lift+Gram state (between begin and finalize) ----
    bool gram_open = false;
    int m = 0, d = 0, p = 0, kind = 0, nk_chunk = 0;
    int MP = 0, KLS = 0, EP = 0, psi_rows = 0, nblk = 0, ntiles = 0;
    int n_pk = 0, n_lf = 0, n_sy = 0, period_len = 0;
    double last_flops = 0.0;
    nk_devbuf zp, zp_in, inv_ls, center, xp[nk::kMaxSlots], yp[nk::kMaxSlots], psi[nk::kMaxSlots], gws, items, counters, tile_of;
    int nslots = 2;
    nk_devbuf gram_err;         // device int: watchdog flag of the fused kernel (nk_gram.cu spin_until_ge)
    int *gram_err_host = nullptr;   // pinned copy, refreshed by nk_gram_finalize; examined at synchronising calls
    bool distinct_in = false;   // input landmarks differ from the output landmarks (regressors.py:133-134 allows injecting them)
    std::vector<int> h_tile_of;

    // ---- dense-stage scratch (grow-only), see nk_dense.cu ----
    nk_devbuf dense[24];
    nk_devbuf dinfo;
    // side stream + events of nk_solve_abc_part: the two regularised systems are independent and each is latency-bound
    // (serial 128x128 diagonal blocks), so the reconstruction solve runs beside the dynamics solve
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace nk {
int gram_watchdog_verdict(nk_handle *h);   // after a stream synchronisation: NK_E_STATE if the fused kernel's watchdog fired
int set_err(nk_handle *h, int code, const std::string &msg);
int check_cuda(nk_handle *h, cudaError_t e, const char *what);
int ensure(nk_handle *h, nk_devbuf &b, size_t bytes);
}  // namespace nk

namespace nk {
// Every entry point runs on the handle's device and leaves the CALLER's current device as it found it (a process may drive
// several GPUs; changing the current device behind the caller's back makes its next allocation land on the wrong GPU).
struct DeviceScope {
    int prev = -1;
    bool changed = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceScope(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); changed = (err == cudaSuccess); }
    }
    ~DeviceScope() { if (changed) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope &) = delete;
    DeviceScope &operator=(const DeviceScope &) = delete;
};
}  // namespace nk

#define NK_ON_DEVICE(h)                                                      \
    nk::DeviceScope _nk_device_scope((h)->device);                           \
    NK_CUDA((h), _nk_device_scope.err)

#define NK_CUDA(h, call)                                                     \
    do {                                                                     \
        int _rc = nk::check_cuda((h), (call), #call);                        \
        if (_rc != NK_OK) return _rc;                                        \
    } while (0)
