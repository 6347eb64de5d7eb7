// ga_engine_internal.h - engine handle shared by ga_engine.cu and ga_host_pipeline.cu (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include "ga_device.cuh"

struct RefEntry { uint32_t* d_ref4 = nullptr; int64_t n = 0; int64_t cap_words = 0; };

// Scratch of one in-flight ga_run: three lanes let the host pipeline overlap consecutive chunks.
struct Lane {
    ga::SessionDesc* d_descs = nullptr; int32_t* d_big_list = nullptr; int32_t* d_large_list = nullptr; int32_t* d_large2_list = nullptr; int64_t cap_sessions = 0;
    int32_t* d_small = nullptr;          // [0] n_big, [1] maxspan, [2..3] tickets, [4..11] fallback reasons, [12] n_large, [13] n_special, [14] second fallback ticket, [15] n_many, [16] n_many_recs, [17] n_kind1, [18] / [19] tickets of the one-CTA / one-warp resolve kernels, [20] sessions the mid one-warp kernel handed to the one-CTA kernel, [21] its ticket
    cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // the fallback kernel runs beside the emission kernel
    uint8_t* d_big_scratch = nullptr;
    // streaming pipeline scratch: scan kernel -> resolve kernel (ga::ScanScratch), resolve -> emission (ga::EmitScratch2)
    uint32_t* d_ent = nullptr; void* d_obs = nullptr; void* d_cnt = nullptr; int64_t cap_items = 0;
    uint8_t* d_kind = nullptr; void* d_edesc = nullptr; void* d_special = nullptr; uint32_t* d_rare_list = nullptr; void* d_edit_keep = nullptr; void* d_many = nullptr; uint32_t* d_many_recs = nullptr; int64_t cap_many = 0; int64_t cap_kind = 0;
    uint32_t* d_germ = nullptr; int64_t cap_germ = 0;
    // CUDA events between the stages of the most recent kTimedRuns runs (ring), recorded on the launching stream so
    // bench.py can read per-launch durations after its timed region without syncing inside it:
    // ev[0] start | scan | ev[1] | resolve (lean + large) | ev[2] | emission | ev[3] | wait for the fallback kernel | ev[4]
    cudaEvent_t ev[5][32] = {};
    int64_t runs = 0;
    // the stream the lane's most recent run was launched on and the end of that run: a run on another stream waits for it
    cudaStream_t last_stream = nullptr; bool used = false; cudaEvent_t ev_done = nullptr;
};
constexpr int kTimedRuns = 32;

constexpr int kLanes = 3;

struct HostSlot;   // ga_host_pipeline.cu

struct ga_engine {
    int device = 0;
    int n_sm = 0;
    std::string err;
    std::map<int, RefEntry> refs;
    Lane lanes[kLanes];
    int64_t big_bytes_per_cta = 0; int big_ctas = 0;
    int32_t big_cols_cap = 1 << 18, big_reads_cap = 1 << 18, big_obs_cap = 1 << 17;
    int64_t launches = 0;
    int last_lane = 0, next_lane = 0;        // ga_run: lane of the most recent run; round-robin cursor when every lane is taken
    int occ_scan = 4, occ_lean = 9, occ_mid = 6, occ_res = 6, occ_rec = 4;
    bool keep_edits = false;              // ga_engine_keep_edits: the resolve kernels keep a copy of every record's edit description   // resident CTAs per SM of the persistent kernels
    HostSlot* slots = nullptr;           // lazily created by ga_run_host
    int64_t last_h2d = 0, last_d2h = 0;
    int64_t* d_fastq_sums = nullptr; int64_t cap_fastq_blocks = 0;   // ga_fastq_layout scratch
};

int ga_fail(ga_engine* e, int code, const char* what, cudaError_t ce = cudaSuccess);
#define GA_CUDA(call) do { cudaError_t _ce = (call); if (_ce != cudaSuccess) return ga_fail(e, GA_ERR_CUDA, #call, _ce); } while (0)

// ga_run on an explicit lane (ga_run itself uses lane 0).
int ga_run_lane(ga_engine* e, int lane, const ga_reads* reads, const ga_sessions* sessions, ga_result* result, cudaStream_t st);
void ga_host_slots_destroy(ga_engine* e);
