// pm_comm.cu -- the residual reduction's second stage fused with the cross-GPU sum of the conformation-sharded objective.
//
// What is replaced (file:line under /root/reference): the reference's "parallel" objective forks a worker pool per call, each
// worker evaluates a slice of the conformations and the parent gathers and sums the slices in Python
// (ParaMol/Objective_function/objective_function.py:268-311, cpu_objective_function.py:67-108).  Here every rank (one process per
// GPU) evaluates its shard; the E(+F) kernel leaves one row of five partial sums per warp, and pm_reduce_rows_kernel
//   (1) adds the rows of its rank in a fixed order (one CTA per parameter vector, one warp per quantity), and
//   (2) -- with a communicator -- is ALSO the all-reduce: the CTA pushes its 8 doubles into the inbox of every peer with
//       peer-to-peer stores over NVLink / NVSwitch, publishes a sequence flag (st.release.sys), waits for the flags of all ranks
//       in its own inbox (ld.acquire.sys) and adds the 8-double rows in rank order.  Every rank therefore holds bit-identical
//       global sums after ONE small kernel -- no NCCL call, no extra launch, nothing on the Python critical path -- and the whole
//       objective call (upload, prepare, evaluate, reduce + all-reduce, download) replays as one CUDA graph on every rank.
//
// Inbox: [2 slots][world][max_vec] entries of 9 x 8 bytes (8 doubles + flag).  The slot alternates with the sequence number; a
// rank can run at most one communicating call ahead of its slowest peer (it needs that peer's flag of call k+1, which the peer
// publishes only after its call-k kernel has finished reading), so two slots suffice.  The sequence number lives in device
// memory (kernel arguments are frozen inside a CUDA graph) and is advanced by the last CTA of each launch.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/paramol_b200.h"
#include "pm_internal.hpp"

#define PMC_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) return pm_fail_msg(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define PMC_ENTRY 9                 // doubles per inbox entry: 8 values + 1 flag
#define PMC_MAX_WORLD 16
#define PMC_THREADS 160             // 5 warps: one per reduced quantity
#define PMC_TIMEOUT_CYCLES 12000000000LL  // ~6 s at 1.97 GHz

struct pm_comm_s {
    int device, rank, world, max_vec;
    bool connected;
    unsigned char *inbox;               // this rank's inbox (device memory, exported through CUDA IPC)
    size_t inbox_bytes;
    void *peer_ptr[PMC_MAX_WORLD];      // peers' inboxes mapped into this process (own entry = inbox)
    void **peers_dev;                   // the same table in device memory
    unsigned long long *state;          // device: [0] sequence number, [1] CTAs finished in the current launch, [2] timeouts
};

struct PmCommArgs {
    void **peers;
    unsigned long long *state;
    int rank, world, max_vec;
};

__device__ __forceinline__ void pmc_st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pmc_ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double pmc_ld_volatile(const double *p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// rows != nullptr: values of vector p = sums over its n_rows rows of 5, [5] = n_conf, [6] = [7] = 0.
// rows == nullptr: values = inout[8 p .. 8 p + 8) (entries >= n_total read as 0): the generic in-place all-reduce.
__global__ void __launch_bounds__(PMC_THREADS)
    pm_reduce_rows_kernel(const double *__restrict__ rows, long n_rows, long n_conf, double *__restrict__ inout, int n_total,
                          PmCommArgs c) {
    __shared__ double s_val[8];
    __shared__ unsigned long long s_seq;
    __shared__ int s_timeout;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, q = tid >> 5;
    if (rows != nullptr) {
        const double *r = rows + (size_t)p * n_rows * 5;
        double t = 0.0;
        for (long b = lane; b < n_rows; b += 32) t += r[b * 5 + q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) s_val[q] = t;
        if (tid == 0) {
            s_val[5] = (double)n_conf;
            s_val[6] = 0.0;
            s_val[7] = 0.0;
        }
    } else if (tid < 8) {
        s_val[tid] = 8 * p + tid < n_total ? inout[8 * p + tid] : 0.0;
    }
    if (tid == 0) {
        s_timeout = 0;
        s_seq = c.world > 1 ? *reinterpret_cast<volatile unsigned long long *>(c.state) + 1ULL : 0ULL;
    }
    __syncthreads();
    if (c.world <= 1) {
        if (tid < 8 && 8 * p + tid < n_total) inout[8 * p + tid] = s_val[tid];
        return;
    }
    const unsigned long long seq = s_seq;
    const size_t slot_off = (size_t)(seq & 1ULL) * c.world * c.max_vec;
    // (1) push this rank's 8 values into every rank's inbox, entry (slot, my rank, p)
    if (tid < 8 * c.world) {
        const int peer = tid >> 3, k = tid & 7;
        double *dst = reinterpret_cast<double *>(c.peers[peer]) + (slot_off + (size_t)c.rank * c.max_vec + p) * PMC_ENTRY;
        dst[k] = s_val[k];
        __threadfence_system();
    }
    __syncthreads();
    // (2) publish, then (3) wait for every rank's entry in the own inbox
    if (tid < c.world) {
        double *dst = reinterpret_cast<double *>(c.peers[tid]) + (slot_off + (size_t)c.rank * c.max_vec + p) * PMC_ENTRY;
        pmc_st_release(reinterpret_cast<unsigned long long *>(dst + 8), seq);
        const double *src = reinterpret_cast<const double *>(c.peers[c.rank]) + (slot_off + (size_t)tid * c.max_vec + p) * PMC_ENTRY;
        const long long t0 = clock64();
        while (pmc_ld_acquire(reinterpret_cast<const unsigned long long *>(src + 8)) != seq) {
            if (clock64() - t0 > PMC_TIMEOUT_CYCLES) {  // a peer never arrived: give up instead of hanging the GPU
                atomicExch(&s_timeout, 1);
                break;
            }
        }
    }
    __syncthreads();
    // (4) the same rank-ordered sum on every rank
    if (tid < 8 && 8 * p + tid < n_total) {
        double total = 0.0;
        for (int r = 0; r < c.world; ++r) {
            const double *src = reinterpret_cast<const double *>(c.peers[c.rank]) + (slot_off + (size_t)r * c.max_vec + p) * PMC_ENTRY;
            total += pmc_ld_volatile(src + tid);
        }
        inout[8 * p + tid] = s_timeout ? nan("") : total;
    }
    if (tid == 0) {
        if (s_timeout) atomicAdd(c.state + 2, 1ULL);
        __threadfence();
        if (atomicAdd(c.state + 1, 1ULL) == (unsigned long long)gridDim.x - 1ULL) {  // last CTA of the launch: advance the sequence
            c.state[1] = 0ULL;
            __threadfence();
            *reinterpret_cast<volatile unsigned long long *>(c.state) = seq;
        }
    }
}

static PmCommArgs pmc_args(pm_comm c) {
    PmCommArgs a;
    a.peers = c ? c->peers_dev : nullptr;
    a.state = c ? c->state : nullptr;
    a.rank = c ? c->rank : 0;
    a.world = (c && c->connected) ? c->world : 1;
    a.max_vec = c ? c->max_vec : 0;
    return a;
}

int pm_comm_max_vec(pm_comm c) { return c ? c->max_vec : 0; }

int pm_reduce_rows(const double *rows_dev, long n_rows, int n_vec, long n_conf, double *partials_dev, pm_comm comm, cudaStream_t stream) {
    if (comm && comm->world > 1 && !comm->connected) return pm_fail_msg("pm_eval_reduce: communicator is not connected");
    pm_reduce_rows_kernel<<<n_vec, PMC_THREADS, 0, stream>>>(rows_dev, n_rows, n_conf, partials_dev, 8 * n_vec, pmc_args(comm));
    PMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int pm_comm_create(int device, int rank, int world, int max_vec, void *handle_out, pm_comm *out) {
    if (!out || !handle_out) return pm_fail_msg("pm_comm_create: null argument");
    if (world < 1 || world > PMC_MAX_WORLD || rank < 0 || rank >= world) return pm_fail_msg("pm_comm_create: bad rank / world size (at most 16 ranks)");
    if (max_vec < 1 || max_vec > 65536) return pm_fail_msg("pm_comm_create: max_vec out of range");
    PMC_CUDA(cudaSetDevice(device));
    pm_comm_s *c = new pm_comm_s();
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->rank = rank;
    c->world = world;
    c->max_vec = max_vec;
    c->connected = world == 1;
    c->inbox_bytes = (size_t)2 * world * max_vec * PMC_ENTRY * sizeof(double);
    auto fail = [&](const std::string &m) {
        if (c->inbox) cudaFree(c->inbox);
        if (c->peers_dev) cudaFree(c->peers_dev);
        if (c->state) cudaFree(c->state);
        delete c;
        return pm_fail_msg(m);
    };
    if (cudaMalloc(&c->inbox, c->inbox_bytes) != cudaSuccess || cudaMemset(c->inbox, 0, c->inbox_bytes) != cudaSuccess ||
        cudaMalloc(&c->peers_dev, sizeof(void *) * PMC_MAX_WORLD) != cudaSuccess ||
        cudaMalloc(&c->state, sizeof(unsigned long long) * 4) != cudaSuccess ||
        cudaMemset(c->state, 0, sizeof(unsigned long long) * 4) != cudaSuccess)
        return fail(std::string("pm_comm_create: ") + cudaGetErrorString(cudaGetLastError()));
    cudaIpcMemHandle_t hd;
    memset(&hd, 0, sizeof(hd));
    if (world > 1 && cudaIpcGetMemHandle(&hd, c->inbox) != cudaSuccess)
        return fail(std::string("pm_comm_create: cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError()));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle_out, &hd, 64);
    c->peer_ptr[rank] = c->inbox;
    if (world == 1 && cudaMemcpy(c->peers_dev, c->peer_ptr, sizeof(void *) * PMC_MAX_WORLD, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail("pm_comm_create: cudaMemcpy failed");
    PMC_CUDA(cudaDeviceSynchronize());
    *out = c;
    return 0;
}

extern "C" int pm_comm_connect(pm_comm c, const void *all_handles) {
    if (!c || !all_handles) return pm_fail_msg("pm_comm_connect: null argument");
    PMC_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const unsigned char *)all_handles + 64 * (size_t)r, 64);
        void *ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return pm_fail_msg(std::string("pm_comm_connect: cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        }
        c->peer_ptr[r] = ptr;
    }
    PMC_CUDA(cudaMemcpy(c->peers_dev, c->peer_ptr, sizeof(void *) * PMC_MAX_WORLD, cudaMemcpyHostToDevice));
    PMC_CUDA(cudaDeviceSynchronize());
    c->connected = true;
    return 0;
}

extern "C" int pm_comm_allreduce(pm_comm c, double *buf_dev, int n, void *stream) {
    if (!c || !buf_dev) return pm_fail_msg("pm_comm_allreduce: null argument");
    if (n < 1 || n > 8 * c->max_vec) return pm_fail_msg("pm_comm_allreduce: buffer larger than 8 * max_vec doubles");
    if (c->world > 1 && !c->connected) return pm_fail_msg("pm_comm_allreduce: communicator is not connected");
    PMC_CUDA(cudaSetDevice(c->device));
    pm_reduce_rows_kernel<<<(n + 7) / 8, PMC_THREADS, 0, (cudaStream_t)stream>>>(nullptr, 0, 0, buf_dev, n, pmc_args(c));
    PMC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" long pm_comm_timeouts(pm_comm c) {
    if (!c) return -1;
    unsigned long long v[4] = {0, 0, 0, 0};
    cudaSetDevice(c->device);
    if (cudaMemcpy(v, c->state, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (long)v[2];
}

extern "C" int pm_comm_destroy(pm_comm c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peer_ptr[r]) cudaIpcCloseMemHandle(c->peer_ptr[r]);
    cudaFree(c->inbox);
    cudaFree(c->peers_dev);
    cudaFree(c->state);
    delete c;
    return 0;
}
