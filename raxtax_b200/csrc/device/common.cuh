// common.cuh -- shared declarations of the sm_100a device library (libraxtax_b200.so).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "raxtax_b200.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

namespace rtx {

constexpr int kWarp = 32;
constexpr u32 kFullMask = 0xFFFFFFFFu;

// Bit-row geometry: one row = one 8-mer, one bit per reference of the shard.  Rows are padded to a whole
// number of 128-word (4096-reference, 512-byte) warp tiles so that every vector load is aligned and in range.
constexpr u32 kRowAlignWords = 128;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;  // a little slack so that slowly growing batches do not realloc
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

// ---- device-side views handed to the kernels ----------------------------------------------------------

struct IndexView {
    const u32* bitrows;   // [n_rows * row_words], row 0 is all-zero
    const u32* rowmap;    // [65536] k-mer -> row id (0 = k-mer absent from this shard)
    const u32* present;   // [2048] bitmap of k-mers with rowmap != 0
    const u64* csr_off;   // [65537]  (only with RTX_OPT_KEEP_CSR)
    const u32* csr_ids;   // [nnz]
    u32 row_words;        // words per bit row (multiple of kRowAlignWords)
    u64 n_refs;           // global N (Tree.num_tips)
    u64 shard_begin;      // first reference id of this shard
    u64 shard_refs;       // references held here
    u64 n_pad;            // row_words * 32: stride of the per-query count vectors
    // lineage tree (Inner / Taxon nodes; children of node i are child_first[i] .. +child_count[i])
    const u32* node_lo;
    const u32* node_hi;
    const u8* node_type;
    const u32* child_first;
    const u32* child_count;
    u32 n_nodes;
    u32 max_levels;
    const u8* ref_levels;  // [n_refs]
    const double* lnfact;  // ln(n!) table
    u32 lnfact_len;
};

struct BatchView {
    u32 n_queries;
    const u64* seq_off;
    const u8* codes;
    const u32* exact_off;  // may be null
    const u32* exact_ids;
    u32 flags;
    u32 kstride;  // stride of kmers / rows per query (multiple of 16)
    u32 hstride;  // stride of hist per query ( >= max K + 1 )
    u16* K;       // [n_queries]
    u16* kmers;   // [n_queries * kstride]
    u32* rows;    // [n_queries * kstride]  bit-row ids of the query's k-mers present in the shard, zero padded
    u32* nrows;   // [n_queries] padded to a multiple of 16
    u32* hist;    // [n_queries * hstride]
};

struct ResultPool {
    u32* first_ref;
    u8* n_levels;
    double* conf;  // [cap * max_levels]
    double* local;
    unsigned long long* used;  // device counter
    u64 cap;
    // per query
    u32* res_off;
    u32* res_cnt;
    double* global_sig;
    int* status;
};

enum QueryStatus : int { kQOk = 0, kQProbSumZero = 1, kQTooManyResults = 2, kQEmptyResult = 3, kQPoolOverflow = 4, kQWalkRetry = 5 };

}  // namespace rtx
