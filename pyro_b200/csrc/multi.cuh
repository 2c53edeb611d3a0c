// Single-process multi-GPU value iteration: ONE host thread drives n slab handles (SURVEY.md 8b "one host thread
// driving n devices ... no MPI/launcher"), so a plain pyro script uses every visible GPU by constructing the planner.
//
// The grid is cut into n_parts slabs over axis 0 (the partition of pyro_b200/distributed.py: balanced, thickness
// differs by at most one plane); part i lives on devices[i] and holds its slab plus the halo its backups can read.
// Within one process the halo planes need no NCCL: after a part's boundary planes are written, cudaMemcpyPeerAsync
// stores them straight into the neighbouring parts' buffers (NVLink peer copies when peer access is available, a
// plain device copy when two parts share a GPU — which is how the slab/halo logic is tested on a 1-GPU box), on a
// side stream under the interior planes.  Per sweep and part:
//     comm stream:  [wait: own previous sweep, neighbours' previous sweep]  boundary-lo, boundary-hi kernels,
//                   peer copies of those planes into the neighbours' halos, event "halo sent"
//     main stream:  interior kernel, [wait: comm stream] statistics fold, event "sweep done";
//                   the NEXT sweep's kernels wait for the neighbours' "halo sent" events.
// The neighbours' previous-sweep wait orders the copy after every kernel that still reads the buffer it overwrites.
// Included by pyrodp.cu (uses the handle internals).
#pragma once

struct pdp_multi {
    std::vector<pdp_handle*> part;
    std::vector<int> device;
    std::vector<cudaEvent_t> ev_done, ev_sent, ev_boundary;   // per part: sweep finished / halos delivered / boundary planes written
    std::vector<double*> dstat;       // per part: [cap][3] folded statistics {jmax, dmax, -dmin} of the enqueued sweeps
    int stats_cap = 0, enqueued = 0;
    int n0 = 0;
    long long N = 0, plane = 0;
    int overlap = 1;
    std::string err;
};

static int mfail(pdp_multi* m, int code, const std::string& msg) {
    if (m) m->err = msg;
    g_err = msg;
    return code;
}

#define MULTI_TRY(m, expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) return mfail(m, PDP_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define PART_TRY(m, expr)                                                                           \
    do {                                                                                            \
        int _rc = (expr);                                                                           \
        if (_rc != PDP_OK) return mfail(m, _rc, g_err);                                             \
    } while (0)

extern "C" int pdp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* pdp_multi_last_error(const pdp_multi* m) { return m ? m->err.c_str() : g_err.c_str(); }

extern "C" int pdp_multi_destroy(pdp_multi* m) {
    if (!m) return PDP_OK;
    for (size_t i = 0; i < m->part.size(); ++i) {
        if (!m->part[i]) continue;
        cudaSetDevice(m->device[i]);
        cudaStreamSynchronize(m->part[i]->stream);
        if (m->part[i]->comm_stream) cudaStreamSynchronize(m->part[i]->comm_stream);
    }
    for (size_t i = 0; i < m->part.size(); ++i) {
        if (i < m->device.size()) cudaSetDevice(m->device[i]);
        if (i < m->ev_done.size() && m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
        if (i < m->ev_sent.size() && m->ev_sent[i]) cudaEventDestroy(m->ev_sent[i]);
        if (i < m->ev_boundary.size() && m->ev_boundary[i]) cudaEventDestroy(m->ev_boundary[i]);
        if (i < m->dstat.size() && m->dstat[i]) cudaFree(m->dstat[i]);
        pdp_destroy(m->part[i]);
    }
    delete m;
    return PDP_OK;
}

// p: the whole-grid descriptor (its slab fields are ignored).  devices[i] = CUDA device of part i; NULL = parts round-robin
// over all visible devices.  A device may appear several times.
extern "C" int pdp_multi_create(const pdp_problem* p, int32_t n_parts, const int32_t* devices, pdp_multi** out) {
    if (!p || !out || n_parts < 1) return mfail(nullptr, PDP_EINVAL, "pdp_multi_create: bad argument");
    *out = nullptr;
    if (p->system_id == PDP_SYS_LUT) return mfail(nullptr, PDP_ENOTSUP, "pdp_multi_create: LUT mode has no a-priori halo; use one handle");
    if (n_parts > p->dims[0]) return mfail(nullptr, PDP_EINVAL, "pdp_multi_create: more parts than axis-0 planes");
    const int ndev = pdp_device_count();
    if (ndev == 0) return mfail(nullptr, PDP_ECUDA, "pdp_multi_create: no usable CUDA device; this engine has no CPU fallback");
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    pdp_multi* m = new pdp_multi();
    auto bail = [&](int code) { std::string e = g_err; pdp_multi_destroy(m); cudaSetDevice(prev_dev); g_err = e; return code; };
    m->n0 = p->dims[0];
    m->N = 1;
    for (int d = 0; d < p->n; ++d) m->N *= p->dims[d];
    m->plane = m->N / m->n0;
    if (const char* env = getenv("PYRODP_MULTI_OVERLAP")) m->overlap = atoi(env) != 0;
    for (int i = 0; i < n_parts; ++i) {
        const int dev = devices ? devices[i] : i % ndev;
        if (dev < 0 || dev >= ndev) { g_err = "pdp_multi_create: device index out of range"; return bail(PDP_EINVAL); }
        m->device.push_back(dev);
    }
    // peer access between the devices of neighbouring parts (ignored where unsupported: the copies are staged then)
    for (int i = 0; i + 1 < n_parts; ++i) {
        const int a = m->device[i], b = m->device[i + 1];
        if (a == b) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) { cudaSetDevice(a); if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) cudaGetLastError(); }
        if (cudaDeviceCanAccessPeer(&can, b, a) == cudaSuccess && can) { cudaSetDevice(b); if (cudaDeviceEnablePeerAccess(a, 0) != cudaSuccess) cudaGetLastError(); }
    }
    m->part.assign(n_parts, nullptr);
    m->ev_done.assign(n_parts, nullptr); m->ev_sent.assign(n_parts, nullptr); m->ev_boundary.assign(n_parts, nullptr);
    m->dstat.assign(n_parts, nullptr);
    m->stats_cap = 256;
    for (int i = 0; i < n_parts; ++i) {
        pdp_problem q = *p;
        q.slab_begin = (int32_t)((long long)i * m->n0 / n_parts);
        q.slab_end = (int32_t)((long long)(i + 1) * m->n0 / n_parts);
        q.alloc_planes = 0;
        if (cudaSetDevice(m->device[i]) != cudaSuccess) { g_err = "pdp_multi_create: cudaSetDevice failed"; return bail(PDP_ECUDA); }
        int rc = pdp_create(&q, &m->part[i]);
        if (rc != PDP_OK) return bail(rc);
        pdp_handle* h = m->part[i];
        if (n_parts > 1 && h->slab_end - h->slab_begin < std::max(h->halo_lo, h->halo_hi)) {
            g_err = "pdp_multi_create: the halo (" + std::to_string(h->halo_lo) + "+" + std::to_string(h->halo_hi) +
                    " planes) is wider than a slab; use fewer parts";
            return bail(PDP_EINVAL);
        }
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        bool ok = cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&m->ev_sent[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&m->ev_boundary[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&m->dstat[i], (size_t)m->stats_cap * 3 * sizeof(double)) == cudaSuccess;
        // events are waited on before anything was recorded in the first sweep: record them once now
        ok = ok && cudaEventRecord(m->ev_done[i], h->stream) == cudaSuccess && cudaEventRecord(m->ev_sent[i], h->comm_stream) == cudaSuccess;
        if (!ok) { g_err = std::string("pdp_multi_create: stream / event setup failed: ") + cudaGetErrorString(cudaGetLastError()); return bail(PDP_ECUDA); }
    }
    cudaSetDevice(prev_dev);
    *out = m;
    return PDP_OK;
}

extern "C" int32_t pdp_multi_parts(const pdp_multi* m) { return m ? (int32_t)m->part.size() : 0; }
extern "C" pdp_handle* pdp_multi_part(const pdp_multi* m, int32_t i) { return (m && i >= 0 && i < (int)m->part.size()) ? m->part[i] : nullptr; }
extern "C" int32_t pdp_multi_part_device(const pdp_multi* m, int32_t i) { return (m && i >= 0 && i < (int)m->device.size()) ? m->device[i] : -1; }
extern "C" int64_t pdp_multi_launch_count(const pdp_multi* m) {
    long long n = 0;
    if (m) for (pdp_handle* h : m->part) n += h->launches;
    return n;
}

// copy planes [p0, p1) of buffer `which` of part src into the same planes of part dst, on stream st of src's device
static int multi_copy_planes(pdp_multi* m, int src, int dst, int which, int p0, int p1, cudaStream_t st) {
    if (p1 <= p0) return PDP_OK;
    pdp_handle *hs = m->part[src], *hd = m->part[dst];
    const double* from = hs->dJ[which] + (long long)(p0 - hs->alloc_begin) * m->plane;
    double* to = hd->dJ[which] + (long long)(p0 - hd->alloc_begin) * m->plane;
    const size_t bytes = (size_t)(p1 - p0) * m->plane * sizeof(double);
    if (m->device[src] == m->device[dst]) MULTI_TRY(m, cudaMemcpyAsync(to, from, bytes, cudaMemcpyDeviceToDevice, st));
    else MULTI_TRY(m, cudaMemcpyPeerAsync(to, m->device[dst], from, m->device[src], bytes, st));
    return PDP_OK;
}

// deliver the slab-boundary planes of buffer `which` of part i to its neighbours' halos (stream st on part i's device)
static int multi_send_halos(pdp_multi* m, int i, int which, cudaStream_t st) {
    pdp_handle* h = m->part[i];
    const int W = (int)m->part.size();
    if (i > 0) {        // part i-1 reads my lowest halo_hi planes
        pdp_handle* nb = m->part[i - 1];
        PART_TRY(m, multi_copy_planes(m, i, i - 1, which, h->slab_begin, std::min(h->slab_begin + nb->halo_hi, nb->alloc_end), st));
    }
    if (i + 1 < W) {    // part i+1 reads my highest halo_lo planes
        pdp_handle* nb = m->part[i + 1];
        PART_TRY(m, multi_copy_planes(m, i, i + 1, which, std::max(h->slab_end - nb->halo_lo, nb->alloc_begin), h->slab_end, st));
    }
    return PDP_OK;
}

extern "C" int pdp_multi_eval_terminal_cost(pdp_multi* m) {
    if (!m) return mfail(nullptr, PDP_EINVAL, "null handle");
    for (pdp_handle* h : m->part) PART_TRY(m, pdp_eval_terminal_cost(h));   // evaluated on slab + halo: no exchange needed
    return PDP_OK;
}

extern "C" int pdp_multi_set_J(pdp_multi* m, const double* J_full) {
    if (!m || !J_full) return mfail(m, PDP_EINVAL, "pdp_multi_set_J: null argument");
    for (pdp_handle* h : m->part) PART_TRY(m, pdp_set_J(h, J_full));
    return PDP_OK;
}

// which: 0 = J, 1 = J_next, 2 = pi; out = the full (N,) array
extern "C" int pdp_multi_get(pdp_multi* m, int32_t which, void* out_full) {
    if (!m || !out_full) return mfail(m, PDP_EINVAL, "pdp_multi_get: null argument");
    if (m->enqueued) return mfail(m, PDP_ESTATE, "pdp_multi_get: collect the enqueued sweeps first");
    for (pdp_handle* h : m->part) {
        const long long off = h->P.slab_node_begin;
        int rc = which == 0 ? pdp_get_J(h, (double*)out_full + off) : which == 1 ? pdp_get_J_next(h, (double*)out_full + off)
                 : which == 2 ? pdp_get_pi(h, (int64_t*)out_full + off) : PDP_EINVAL;
        if (rc != PDP_OK) return mfail(m, rc, which > 2 || which < 0 ? "pdp_multi_get: which must be 0, 1 or 2" : g_err);
    }
    return PDP_OK;
}

extern "C" int pdp_multi_get_input_from_policy(pdp_multi* m, int32_t k, double* uk_full) {
    if (!m || !uk_full) return mfail(m, PDP_EINVAL, "pdp_multi_get_input_from_policy: null argument");
    for (pdp_handle* h : m->part) PART_TRY(m, pdp_get_input_from_policy(h, k, uk_full + h->P.slab_node_begin));
    return PDP_OK;
}

static int multi_sync_all(pdp_multi* m) {
    for (size_t i = 0; i < m->part.size(); ++i) {
        MULTI_TRY(m, cudaSetDevice(m->device[i]));
        MULTI_TRY(m, cudaStreamSynchronize(m->part[i]->comm_stream));
        MULTI_TRY(m, cudaStreamSynchronize(m->part[i]->stream));
    }
    return PDP_OK;
}

extern "C" int pdp_multi_clean_infeasible_set(pdp_multi* m, double tol, int64_t default_action) {
    if (!m) return mfail(nullptr, PDP_EINVAL, "null handle");
    for (pdp_handle* h : m->part) PART_TRY(m, pdp_clean_infeasible_set(h, tol, default_action));   // slab nodes ...
    for (size_t i = 0; i < m->part.size(); ++i) {                                                  // ... then the neighbours' halo copies
        MULTI_TRY(m, cudaSetDevice(m->device[i]));
        PART_TRY(m, multi_send_halos(m, (int)i, m->part[i]->cur_idx, m->part[i]->stream));
    }
    return multi_sync_all(m);
}

// Enqueue one sweep of every part (+ halo delivery) without host synchronisation.
extern "C" int pdp_multi_sweep_enqueue(pdp_multi* m) {
    if (!m) return mfail(nullptr, PDP_EINVAL, "null handle");
    const int W = (int)m->part.size();
    for (pdp_handle* h : m->part) {
        if (h->sticky) return mfail(m, h->sticky, h->err);
        if (!h->have_J) return mfail(m, PDP_ESTATE, "pdp_multi_sweep: no cost-to-go yet (pdp_multi_set_J / pdp_multi_eval_terminal_cost)");
    }
    if (m->enqueued >= m->stats_cap) return mfail(m, PDP_ESTATE, "pdp_multi_sweep: too many sweeps enqueued; collect first");
    // phase 1: every part's kernels and outgoing copies.  All waits refer to events recorded in the PREVIOUS sweep (or at
    // creation), so the order in which the parts are visited does not matter.
    for (int i = 0; i < W; ++i) {
        pdp_handle* h = m->part[i];
        MULTI_TRY(m, cudaSetDevice(m->device[i]));
        const int b = h->slab_begin, e = h->slab_end;
        double* sets = h->dstats_sets;
        double* dst = m->dstat[i] + 3 * m->enqueued;
        // this sweep reads J[cur] incl. the halo planes the neighbours delivered in the previous sweep
        for (int nb = i - 1; nb <= i + 1; nb += 2) {
            if (nb < 0 || nb >= W) continue;
            MULTI_TRY(m, cudaStreamWaitEvent(h->stream, m->ev_sent[nb], 0));
            MULTI_TRY(m, cudaStreamWaitEvent(h->comm_stream, m->ev_sent[nb], 0));
            // ... and its outgoing copies overwrite halo planes of the neighbours' J[1-cur], which their previous sweep read
            MULTI_TRY(m, cudaStreamWaitEvent(h->comm_stream, m->ev_done[nb], 0));
        }
        MULTI_TRY(m, cudaStreamWaitEvent(h->comm_stream, m->ev_done[i], 0));      // own previous sweep (it wrote J[cur])
        MULTI_TRY(m, cudaStreamWaitEvent(h->stream, m->ev_sent[i], 0));           // own outgoing copies of the previous sweep
        const int blo = (i > 0) ? m->part[i - 1]->halo_hi : 0;       // planes the lower / upper neighbour needs
        const int bhi = (i + 1 < W) ? m->part[i + 1]->halo_lo : 0;
        if (W > 1 && m->overlap && b + blo < e - bhi) {
            if (blo > 0) PART_TRY(m, launch_planes(h, b, b + blo, 1, sets + 3, h->comm_stream));
            if (bhi > 0) PART_TRY(m, launch_planes(h, e - bhi, e, 2, sets + 6, h->comm_stream));
            PART_TRY(m, multi_send_halos(m, i, 1 - h->cur_idx, h->comm_stream));
            MULTI_TRY(m, cudaEventRecord(m->ev_boundary[i], h->comm_stream));
            PART_TRY(m, launch_planes(h, b + blo, e - bhi, 0, sets));
            MULTI_TRY(m, cudaStreamWaitEvent(h->stream, m->ev_boundary[i], 0));
            // sets 1 / 2 hold stale triples when a boundary is empty: fold only what was launched
            if (blo > 0 && bhi > 0) stats_fold_kernel<<<1, 1, 0, h->stream>>>(sets, 3, dst);
            else if (blo > 0) stats_fold_kernel<<<1, 1, 0, h->stream>>>(sets, 2, dst);
            else { stats_fold2_kernel<<<1, 1, 0, h->stream>>>(sets, sets + 6, dst); }
        } else {
            PART_TRY(m, launch_planes(h, b, e, 0, sets));
            stats_fold_kernel<<<1, 1, 0, h->stream>>>(sets, 1, dst);
            MULTI_TRY(m, cudaEventRecord(m->ev_boundary[i], h->stream));
            MULTI_TRY(m, cudaStreamWaitEvent(h->comm_stream, m->ev_boundary[i], 0));
            PART_TRY(m, multi_send_halos(m, i, 1 - h->cur_idx, h->comm_stream));
        }
        MULTI_TRY(m, cudaGetLastError());
    }
    // phase 2: publish this sweep's events only after every part has enqueued its waits on the previous ones
    for (int i = 0; i < W; ++i) {
        pdp_handle* h = m->part[i];
        MULTI_TRY(m, cudaSetDevice(m->device[i]));
        MULTI_TRY(m, cudaEventRecord(m->ev_sent[i], h->comm_stream));
        MULTI_TRY(m, cudaEventRecord(m->ev_done[i], h->stream));
        h->cur_idx = 1 - h->cur_idx;
    }
    m->enqueued += 1;
    return PDP_OK;
}

extern "C" int pdp_multi_sweep_collect(pdp_multi* m, pdp_stats* stats_out, int32_t max_out, int32_t* n_out) {
    if (!m) return mfail(nullptr, PDP_EINVAL, "null handle");
    const int n = m->enqueued, k = std::min(n, (int)max_out);
    if (n_out) *n_out = k;
    PART_TRY(m, multi_sync_all(m));
    std::vector<double> tmp((size_t)std::max(n, 1) * 3);
    for (size_t i = 0; i < m->part.size(); ++i) {
        if (n == 0) break;
        MULTI_TRY(m, cudaSetDevice(m->device[i]));
        MULTI_TRY(m, cudaMemcpy(tmp.data(), m->dstat[i], (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        for (int s = 0; s < k && stats_out; ++s) {
            pdp_stats v = {tmp[3 * s], tmp[3 * s + 1], -tmp[3 * s + 2]};
            if (i == 0) stats_out[s] = v;
            else {
                stats_out[s].j_max = std::max(stats_out[s].j_max, v.j_max);
                stats_out[s].delta_max = std::max(stats_out[s].delta_max, v.delta_max);
                stats_out[s].delta_min = std::min(stats_out[s].delta_min, v.delta_min);
            }
        }
    }
    m->enqueued = 0;
    return PDP_OK;
}

extern "C" int pdp_multi_sweep(pdp_multi* m, int32_t n_sweeps, pdp_stats* stats_out) {
    if (!m || n_sweeps < 0) return mfail(m, PDP_EINVAL, "pdp_multi_sweep: bad argument");
    if (m->enqueued) return mfail(m, PDP_ESTATE, "pdp_multi_sweep: collect the enqueued sweeps first");
    int done = 0;
    while (done < n_sweeps) {           // batches bounded by the statistics history
        const int batch = std::min(n_sweeps - done, m->stats_cap);
        for (int s = 0; s < batch; ++s) {
            int rc = pdp_multi_sweep_enqueue(m);
            if (rc != PDP_OK) { m->enqueued = 0; return rc; }
        }
        int rc = pdp_multi_sweep_collect(m, stats_out ? stats_out + done : nullptr, batch, nullptr);
        if (rc != PDP_OK) return rc;
        done += batch;
    }
    return PDP_OK;
}
