// vcm_tables: immutable per-layer connectivity on the device + library plumbing.
//
// Replaces the table parsing of LASMConvssw.__init__ (graphVAESSW.py:44-57)
// and removes the per-call mask construction of graphVAESSW.py:111-118: the
// kernels test the sentinel id directly.
#include <stdarg.h>

#include <algorithm>
#include <atomic>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace vcm {

std::string &last_error_slot() {
    static thread_local std::string s;
    return s;
}

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_slot() = buf;
    return code;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        cached = cudaGetDeviceProperties(&p, dev) == cudaSuccess ? p.multiProcessorCount : 148;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace vcm

using vcm::Tables;

extern "C" {

const char *vcm_last_error(void) { return vcm::last_error_slot().c_str(); }
int vcm_version(void) { return 100; }
unsigned long long vcm_launch_count(void) { return vcm::launches(); }

int vcm_tables_create(const int32_t *table, int p_out, int cols, int p_in, vcm_tables **out) {
    VCM_CHECK(out != nullptr, "out is NULL");
    *out = nullptr;
    VCM_CHECK(table != nullptr, "table_host is NULL");
    VCM_CHECK(p_out >= 0 && p_in >= 0, "negative vertex count");
    VCM_CHECK(cols >= 1 && (cols - 1) % 2 == 0, "table must have 1 + 2N columns, got %d", cols);
    VCM_CHECK(!vcm::is_device_ptr(table), "table_host must be a host pointer");
    const int n_max = (cols - 1) / 2;
    VCM_CHECK((int64_t)p_out * std::max(n_max, 1) < (int64_t)INT32_MAX, "table too large for int32 slot ids");

    std::vector<int32_t> idx((size_t)p_out * n_max), last(p_out), perm(p_out);
    std::vector<int32_t> rev_ptr((size_t)p_in + 1, 0);
    int64_t nnz = 0;
    int max_last = 0;
    for (int p = 0; p < p_out; ++p) {
        const int32_t *row = table + (size_t)p * cols;
        int l = 0;
        for (int n = 0; n < n_max; ++n) {
            const int32_t q = row[1 + 2 * n];
            VCM_CHECK(q >= 0 && q <= p_in, "table[%d] neighbour %d = %d outside [0, %d]", p, n, q, p_in);
            idx[(size_t)p * n_max + n] = q;
            if (q != p_in) {
                l = n + 1;
                ++nnz;
                ++rev_ptr[q + 1];
            }
        }
        last[p] = l;
        max_last = std::max(max_last, l);
    }
    // degree buckets: stable sort by `last` descending (counting sort)
    {
        std::vector<int32_t> start(max_last + 2, 0);
        for (int p = 0; p < p_out; ++p) ++start[max_last - last[p] + 1];
        for (int d = 0; d <= max_last; ++d) start[d + 1] += start[d];
        for (int p = 0; p < p_out; ++p) perm[start[max_last - last[p]]++] = p;
    }
    // reverse CSR: slots ascending per input vertex (p ascending, then n)
    int max_rev = 0;
    for (int q = 0; q < p_in; ++q) {
        max_rev = std::max(max_rev, rev_ptr[q + 1]);
        rev_ptr[q + 1] += rev_ptr[q];
    }
    std::vector<int32_t> rev_slot((size_t)nnz);
    {
        std::vector<int32_t> fill(rev_ptr.begin(), rev_ptr.end() - 1);
        for (int p = 0; p < p_out; ++p)
            for (int n = 0; n < n_max; ++n) {
                const int32_t q = idx[(size_t)p * n_max + n];
                if (q != p_in) rev_slot[fill[q]++] = p * n_max + n;
            }
    }

    // processing-order copies: a CTA finds vertex id, trip count and neighbour ids of its work item at
    // addresses that depend on the slot only (no perm -> last -> idx chain of dependent loads)
    std::vector<int32_t> idx_sorted(idx.size()), meta_sorted((size_t)p_out * 2);
    for (int s2 = 0; s2 < p_out; ++s2) {
        const int p = perm[s2];
        meta_sorted[2 * (size_t)s2] = p;
        meta_sorted[2 * (size_t)s2 + 1] = last[p];
        std::copy(idx.begin() + (size_t)p * n_max, idx.begin() + (size_t)(p + 1) * n_max,
                  idx_sorted.begin() + (size_t)s2 * n_max);
    }

    Tables *t = new (std::nothrow) Tables;
    if (!t) return vcm::fail(VCM_ENOMEM, "out of host memory");
    t->p_in = p_in;
    t->p_out = p_out;
    t->n_max = n_max;
    t->max_last = max_last;
    t->max_rev = max_rev;
    t->n_gt16 = 0;
    for (int p = 0; p < p_out; ++p) t->n_gt16 += last[p] > 16;
    t->nnz = nnz;
    cudaGetDevice(&t->device);

    // one allocation, 64-element aligned sub-arrays
    auto pad = [](size_t n) { return (n + 63) / 64 * 64; };
    const size_t o_idx = 0, o_last = o_idx + pad(idx.size()), o_perm = o_last + pad(last.size()),
                 o_ptr = o_perm + pad(perm.size()), o_slot = o_ptr + pad(rev_ptr.size()),
                 o_is = o_slot + pad(rev_slot.size()), o_ms = o_is + pad(idx_sorted.size()),
                 total = o_ms + pad(meta_sorted.size()) + 64;
    cudaError_t e = cudaMalloc(&t->blob, total * sizeof(int32_t));
    if (e != cudaSuccess) {
        delete t;
        return vcm::fail((int)e, "cudaMalloc(%zu) failed: %s", total * sizeof(int32_t), cudaGetErrorString(e));
    }
    t->idx = t->blob + o_idx;
    t->last = t->blob + o_last;
    t->perm = t->blob + o_perm;
    t->rev_ptr = t->blob + o_ptr;
    t->rev_slot = t->blob + o_slot;
    t->idx_sorted = t->blob + o_is;
    t->meta_sorted = t->blob + o_ms;
    auto up = [&](int32_t *dst, const std::vector<int32_t> &src) {
        return src.empty() ? cudaSuccess
                           : cudaMemcpy(dst, src.data(), src.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    };
    if ((e = up(t->idx, idx)) != cudaSuccess || (e = up(t->last, last)) != cudaSuccess ||
        (e = up(t->perm, perm)) != cudaSuccess || (e = up(t->rev_ptr, rev_ptr)) != cudaSuccess ||
        (e = up(t->rev_slot, rev_slot)) != cudaSuccess || (e = up(t->idx_sorted, idx_sorted)) != cudaSuccess ||
        (e = up(t->meta_sorted, meta_sorted)) != cudaSuccess) {
        cudaFree(t->blob);
        delete t;
        return vcm::fail((int)e, "table upload failed: %s", cudaGetErrorString(e));
    }
    *out = reinterpret_cast<vcm_tables *>(t);
    return VCM_OK;
}

void vcm_tables_destroy(vcm_tables *h) {
    Tables *t = reinterpret_cast<Tables *>(h);
    if (!t) return;
    if (t->blob) cudaFree(t->blob);
    delete t;
}

int vcm_tables_get_info(const vcm_tables *h, vcm_tables_info *info) {
    VCM_CHECK(h && info, "null argument");
    const Tables *t = vcm::as_tables(h);
    info->p_in = t->p_in;
    info->p_out = t->p_out;
    info->n_max = t->n_max;
    info->max_last = t->max_last;
    info->nnz = t->nnz;
    info->max_rev = t->max_rev;
    info->device = t->device;
    return VCM_OK;
}

static const int32_t *pick(const Tables *t, int which, size_t *n) {
    switch (which) {
        case VCM_TAB_IDX: *n = (size_t)t->p_out * t->n_max; return t->idx;
        case VCM_TAB_LAST: *n = t->p_out; return t->last;
        case VCM_TAB_PERM: *n = t->p_out; return t->perm;
        case VCM_TAB_REV_PTR: *n = (size_t)t->p_in + 1; return t->rev_ptr;
        case VCM_TAB_REV_SLOT: *n = (size_t)t->nnz; return t->rev_slot;
        case VCM_TAB_IDX_SORTED: *n = (size_t)t->p_out * t->n_max; return t->idx_sorted;
        case VCM_TAB_META_SORTED: *n = (size_t)t->p_out * 2; return t->meta_sorted;
    }
    *n = 0;
    return nullptr;
}

int vcm_tables_export(const vcm_tables *h, int which, int32_t *dst, size_t capacity) {
    VCM_CHECK(h && dst, "null argument");
    size_t n;
    const int32_t *src = pick(vcm::as_tables(h), which, &n);
    VCM_CHECK(src != nullptr, "unknown table id %d", which);
    VCM_CHECK(capacity >= n, "buffer holds %zu elements, table has %zu", capacity, n);
    if (n) VCM_CUDA(cudaMemcpy(dst, src, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return VCM_OK;
}

const int32_t *vcm_tables_device_ptr(const vcm_tables *h, int which) {
    if (!h) return nullptr;
    size_t n;
    return pick(vcm::as_tables(h), which, &n);
}

}  // extern "C"
