// comm.cu -- the path's only exchange step, inside the library: the final gather of every rank's [nq, k] results
// (the reference's single strategy is rayon over queries in one process,
// index/hnsw/index/batch.rs:159-197 -- here the queries of a batch are sharded over the GPUs of one NVSwitch box).
//
// No collective library on the data path.  Every rank owns a gather window [world][nq][k] (ids, distances, counts) in
// its own HBM and maps its peers' windows through CUDA IPC.  The search kernel's epilogue (hnsw_search.cuh) stores each
// finished query's top-k into slot `rank` of EVERY window -- its own and, as posted stores over NVLink, the peers' --
// so the transfer rides under the batch's remaining queries instead of following it.  What is left after the kernel is
// a flag exchange: comm_signal_wait_kernel tells every peer "my slot of epoch e is complete" and spins until every
// peer has said the same.  Windows are double buffered by epoch parity: a rank can run at most one epoch ahead of a
// peer (it needs the peer's flag of epoch e to leave epoch e), so it never overwrites a slot the peer may still read.
#include <algorithm>

#include "hnsw_search.cuh"

struct veles_comm {
    int rank = 0, world = 1;
    uint32_t nq = 0, k = 0;
    uint64_t epoch = 0;
    size_t ids_off[2] = {0, 0}, dist_off[2] = {0, 0}, cnt_off[2] = {0, 0};
    veles::DevBuf window, flags, err;
    uint8_t* peer_window[8] = {nullptr};
    uint32_t* peer_flags[8] = {nullptr};
    bool connected = false;
};

namespace veles {

// One warp.  Lane r < world, r != rank: publish "epoch e of rank `rank` is in your window", then wait for the same
// from r.  The search kernel's stores were issued by an earlier kernel on this stream, so they have been performed
// (kernel boundary) before this kernel's system-scope fence and flag store.  A bounded spin: a lost peer turns into an
// error flag, never into a hung GPU.
__global__ void __launch_bounds__(32) comm_signal_wait_kernel(uint32_t rank, uint32_t world, uint32_t epoch,
                                                              uint32_t* const* __restrict__ peer_flags_tab,
                                                              volatile uint32_t* my_flags, uint32_t* err,
                                                              long long timeout_cycles) {
    const uint32_t r = threadIdx.x;
    if (r >= world || r == rank) return;
    __threadfence_system();
    volatile uint32_t* dst = peer_flags_tab[r] + rank;
    *dst = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    while ((int32_t)(my_flags[r] - epoch) < 0) {
        if (clock64() - t0 > timeout_cycles) {
            atomicExch(err, 1u + r);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace veles

using namespace veles;

extern "C" {

int32_t veles_comm_handle_bytes(void) { return (int32_t)(2 * sizeof(cudaIpcMemHandle_t)); }

int32_t veles_comm_create(int32_t rank, int32_t world, uint32_t nq_per_rank, uint32_t k, veles_comm_t** out, void* handle_out) {
    VELES_REQUIRE(out != nullptr && handle_out != nullptr, "NULL argument");
    *out = nullptr;
    VELES_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "rank %d / world %d: one NVSwitch box holds 1..8 GPUs", rank,
                  world);
    VELES_REQUIRE(nq_per_rank >= 1 && k >= 1, "empty window");
    std::unique_ptr<veles_comm> c(new veles_comm());
    c->rank = rank;
    c->world = world;
    c->nq = nq_per_rank;
    c->k = k;
    const size_t per = (size_t)world * nq_per_rank;
    const size_t ids_b = round_up((uint32_t)std::min<size_t>(per * k * 4, 0xffffff00u), 256);
    VELES_REQUIRE(per * k * 4 < 0xffffff00u, "gather window too large");
    const size_t cnt_b = round_up((uint32_t)(per * 4), 256);
    size_t off = 0;
    for (int b = 0; b < 2; ++b) {
        c->ids_off[b] = off;
        off += ids_b;
        c->dist_off[b] = off;
        off += ids_b;
        c->cnt_off[b] = off;
        off += cnt_b;
    }
    VELES_TRY(c->window.alloc(off));
    VELES_TRY(c->flags.alloc(256));
    VELES_TRY(c->err.alloc(16));
    VELES_CUDA(cudaMemset(c->window.p, 0xff, off));
    VELES_CUDA(cudaMemset(c->flags.p, 0, 256));
    VELES_CUDA(cudaMemset(c->err.p, 0, 16));
    cudaIpcMemHandle_t h[2];
    VELES_CUDA(cudaIpcGetMemHandle(&h[0], c->window.p));
    VELES_CUDA(cudaIpcGetMemHandle(&h[1], c->flags.p));
    std::memcpy(handle_out, h, sizeof(h));
    c->peer_window[rank] = c->window.as<uint8_t>();
    c->peer_flags[rank] = c->flags.as<uint32_t>();
    *out = c.release();
    return VELES_OK;
}

// all_handles: world x veles_comm_handle_bytes(), rank r's blob at position r (exchanged by the host over whatever
// channel it already has; every rank must have called veles_comm_create with the same nq_per_rank and k)
int32_t veles_comm_connect(veles_comm_t* c, const void* all_handles) {
    VELES_REQUIRE(c != nullptr && all_handles != nullptr, "NULL argument");
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        void* w = nullptr;
        void* f = nullptr;
        VELES_CUDA(cudaIpcOpenMemHandle(&w, h[2 * r], cudaIpcMemLazyEnablePeerAccess));
        VELES_CUDA(cudaIpcOpenMemHandle(&f, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
        c->peer_window[r] = static_cast<uint8_t*>(w);
        c->peer_flags[r] = static_cast<uint32_t*>(f);
    }
    // device-side table of the peers' flag arrays, kept at the end of the local flag buffer
    VELES_CUDA(cudaMemcpy(c->flags.as<uint8_t>() + 128, c->peer_flags, sizeof(uint32_t*) * 8, cudaMemcpyHostToDevice));
    c->connected = true;
    return VELES_OK;
}

// NativeHnsw::search for this rank's `nq` queries (as veles_search_batch_d) with the gather fused in: results land in
// slot `rank` of every rank's window; when the work enqueued here has run, the local window of this epoch holds every
// rank's results.  Collective: every rank calls it, in the same order, with the same nq and k.  Only enqueues.
int32_t veles_search_batch_gather_d(const veles_index_t* idx, veles_comm_t* c, const float* queries_d, uint32_t nq, uint32_t k,
                                    uint32_t ef, void* stream, void* mid_event) {
    VELES_REQUIRE(idx != nullptr && c != nullptr, "NULL argument");
    VELES_REQUIRE(c->connected || c->world == 1, "veles_comm_connect has not been called");
    VELES_REQUIRE(nq == c->nq && k == c->k, "the window was created for nq=%u, k=%u; got %u, %u", c->nq, c->k, nq, k);
    VELES_REQUIRE(queries_d != nullptr, "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    c->epoch += 1;
    const int b = (int)(c->epoch & 1);
    const size_t slot_e = (size_t)c->rank * nq * k * 4, slot_c = (size_t)c->rank * nq * 4;
    PeerOut po;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        po.ids[po.n] = reinterpret_cast<uint32_t*>(c->peer_window[r] + c->ids_off[b] + slot_e);
        po.dist[po.n] = reinterpret_cast<float*>(c->peer_window[r] + c->dist_off[b] + slot_e);
        po.cnt[po.n] = reinterpret_cast<uint32_t*>(c->peer_window[r] + c->cnt_off[b] + slot_c);
        po.n++;
    }
    uint8_t* w = c->window.as<uint8_t>();
    {
        std::lock_guard<std::mutex> g(idx->mu);
        SearchCtx* ctx = nullptr;
        VELES_TRY(acquire_ctx(idx, st, false, &ctx));
        VELES_TRY(launch_search(idx, idx->view(), ctx, queries_d, nq, k, ef, reinterpret_cast<uint32_t*>(w + c->ids_off[b] + slot_e),
                                reinterpret_cast<float*>(w + c->dist_off[b] + slot_e),
                                reinterpret_cast<uint32_t*>(w + c->cnt_off[b] + slot_c), nullptr, st, nullptr, &po));
    }
    if (mid_event) VELES_CUDA(cudaEventRecord((cudaEvent_t)mid_event, st));
    if (c->world > 1) {
        comm_signal_wait_kernel<<<1, 32, 0, st>>>((uint32_t)c->rank, (uint32_t)c->world, (uint32_t)c->epoch,
                                                  reinterpret_cast<uint32_t* const*>(c->flags.as<uint8_t>() + 128),
                                                  c->flags.as<uint32_t>(), c->err.as<uint32_t>(), 4000000000ll);
        count_launch();
        VELES_CUDA(cudaGetLastError());
    }
    return VELES_OK;
}

// The local window of the most recent epoch: [world * nq * k] ids and distances, [world * nq] counts, rank-major --
// i.e. the whole batch in query order.  Valid once the stream of the last veles_search_batch_gather_d call has run.
int32_t veles_comm_window(veles_comm_t* c, uint32_t** ids_d, float** dist_d, uint32_t** counts_d) {
    VELES_REQUIRE(c != nullptr, "comm is NULL");
    const int b = (int)(c->epoch & 1);
    uint8_t* w = c->window.as<uint8_t>();
    if (ids_d) *ids_d = reinterpret_cast<uint32_t*>(w + c->ids_off[b]);
    if (dist_d) *dist_d = reinterpret_cast<float*>(w + c->dist_off[b]);
    if (counts_d) *counts_d = reinterpret_cast<uint32_t*>(w + c->cnt_off[b]);
    return VELES_OK;
}

// waits for `stream`; VELES_ERR_CUDA if a peer never signalled (its rank is in the message)
int32_t veles_comm_status(veles_comm_t* c, void* stream) {
    VELES_REQUIRE(c != nullptr, "comm is NULL");
    VELES_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    uint32_t h = 0;
    VELES_CUDA(cudaMemcpy(&h, c->err.p, 4, cudaMemcpyDeviceToHost));
    if (h != 0) {
        set_error("gather: rank %u never signalled epoch %llu (timeout)", h - 1, (unsigned long long)c->epoch);
        return VELES_ERR_CUDA;
    }
    return VELES_OK;
}

int32_t veles_comm_destroy(veles_comm_t* c) {
    if (!c) return VELES_OK;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        if (c->peer_window[r]) cudaIpcCloseMemHandle(c->peer_window[r]);
        if (c->peer_flags[r]) cudaIpcCloseMemHandle(c->peer_flags[r]);
    }
    delete c;
    return VELES_OK;
}

}  // extern "C"
