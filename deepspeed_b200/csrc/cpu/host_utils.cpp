// Host-side helpers for ragged (continuous-batching) inference and block-sparse attention.
//
// Role parity: reference inference/v2/kernels/ragged_ops/atom_builder (N9b host part),
// inference/v2/ragged/csrc/fast_host_buffer.cu (N9d, page-locked staging) and
// csrc/sparse_attention/utils.cpp `sdd_segment` (N15).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define DSB_EXPORT extern "C" __attribute__((visibility("default")))

// Attention "atoms": a unit of work = (sequence, block of q tokens, range of kv blocks).  For each
// in-flight sequence with `n_new` new tokens and `n_seen` cached tokens, emit one atom per
// q-block of `q_block` tokens covering kv blocks [0, ceil((n_seen + q_end)/kv_block)).
// Output rows: {seq_idx, q_start (global token index), q_len, kv_blocks, total_kv_len, block_table_offset}.
DSB_EXPORT int64_t dsb_build_atoms(const int32_t* seq_new_tokens, const int32_t* seq_seen_tokens,
                                   const int32_t* seq_token_start, const int32_t* seq_block_table_off, int32_t n_seqs,
                                   int32_t q_block, int32_t kv_block, int32_t* atoms_out, int64_t max_atoms)
{
    int64_t n = 0;
    for (int32_t s = 0; s < n_seqs; ++s) {
        const int32_t n_new = seq_new_tokens[s], n_seen = seq_seen_tokens[s];
        for (int32_t q0 = 0; q0 < n_new; q0 += q_block) {
            if (n >= max_atoms) return -1;
            const int32_t qlen = std::min(q_block, n_new - q0);
            const int32_t kv_len = n_seen + q0 + qlen;
            int32_t* a = atoms_out + n * 6;
            a[0] = s;
            a[1] = seq_token_start[s] + q0;
            a[2] = qlen;
            a[3] = (kv_len + kv_block - 1) / kv_block;
            a[4] = kv_len;
            a[5] = seq_block_table_off[s];
            ++n;
        }
    }
    return n;
}

// Block-sparse layout segmentation: given a [H, M, N] 0/1 layout, greedily cover each head's
// non-zero blocks with maximal square segments of size `max_width` down to 1 (used to batch SDD
// work).  Output rows: {head, row, col, width}.  Returns number of segments.
DSB_EXPORT int64_t dsb_sdd_segment(const int32_t* layout, int32_t H, int32_t M, int32_t N, int32_t max_width,
                                   int32_t* out, int64_t max_out)
{
    std::vector<uint8_t> used(static_cast<size_t>(H) * M * N, 0);
    int64_t n = 0;
    for (int32_t w = max_width; w >= 1; w /= 2) {
        for (int32_t h = 0; h < H; ++h) {
            for (int32_t i = 0; i + w <= M; i += w) {
                for (int32_t j = 0; j + w <= N; j += w) {
                    bool ok = true;
                    for (int32_t a = 0; a < w && ok; ++a)
                        for (int32_t b = 0; b < w; ++b) {
                            const size_t idx = (static_cast<size_t>(h) * M + i + a) * N + j + b;
                            if (!layout[idx] || used[idx]) {
                                ok = false;
                                break;
                            }
                        }
                    if (!ok) continue;
                    if (n >= max_out) return -1;
                    for (int32_t a = 0; a < w; ++a)
                        for (int32_t b = 0; b < w; ++b) used[(static_cast<size_t>(h) * M + i + a) * N + j + b] = 1;
                    int32_t* o = out + n * 4;
                    o[0] = h;
                    o[1] = i;
                    o[2] = j;
                    o[3] = w;
                    ++n;
                }
            }
        }
        if (w == 1) break;
    }
    return n;
}

// Free-list block allocator for the paged KV cache (host bookkeeping, O(1) alloc/free).
struct BlockAllocator {
    std::vector<int32_t> next;
    int32_t head;
    int32_t free_blocks;
};

DSB_EXPORT void* dsb_blockalloc_create(int32_t n_blocks)
{
    auto* a = new BlockAllocator();
    a->next.resize(static_cast<size_t>(n_blocks));
    for (int32_t i = 0; i < n_blocks; ++i) a->next[i] = i + 1 < n_blocks ? i + 1 : -1;
    a->head = n_blocks > 0 ? 0 : -1;
    a->free_blocks = n_blocks;
    return a;
}
DSB_EXPORT void dsb_blockalloc_destroy(void* h) { delete static_cast<BlockAllocator*>(h); }
DSB_EXPORT int32_t dsb_blockalloc_free_count(void* h) { return static_cast<BlockAllocator*>(h)->free_blocks; }
DSB_EXPORT int32_t dsb_blockalloc_allocate(void* h, int32_t n, int32_t* out)
{
    auto* a = static_cast<BlockAllocator*>(h);
    if (n > a->free_blocks) return -1;
    for (int32_t i = 0; i < n; ++i) {
        out[i] = a->head;
        a->head = a->next[a->head];
        a->next[out[i]] = -2;  // allocated marker
    }
    a->free_blocks -= n;
    return 0;
}
DSB_EXPORT int32_t dsb_blockalloc_free(void* h, const int32_t* blocks, int32_t n)
{
    auto* a = static_cast<BlockAllocator*>(h);
    for (int32_t i = 0; i < n; ++i) {
        const int32_t b = blocks[i];
        if (b < 0 || b >= static_cast<int32_t>(a->next.size()) || a->next[b] != -2) return -1;  // double free
        a->next[b] = a->head;
        a->head = b;
    }
    a->free_blocks += n;
    return 0;
}
