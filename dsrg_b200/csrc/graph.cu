// Whole device passes as CUDA graphs (sm_100a host side).
//
// The reference runs its refinement per image in a Python loop (pylayers/pylayers/pylayers.py:325-326); here a
// pass over a batch is 40-130 dependent kernel launches (lattice build, T+1 tile launches, 9T blur launches, SRG).
// At the shapes the reference trains on (20 x 21 x 41 x 41, train-s.prototxt:17) and for the chunks of the
// host-buffer pipeline those launches are pure latency (~10 us each for ~1-3 us of work).  All grid sizes and
// pointers of a pass are functions of the call's arguments (every count the kernels need stays on the device),
// so a pass is captured once per distinct argument set and replayed with one cudaGraphLaunch.
#include "common.cuh"

namespace dsrg {

static const int kMaxGraphs = 48;  // per engine; least recently used ones are dropped

static void drop(Engine::GraphRec &r) {
    if (r.exec) cudaGraphExecDestroy(r.exec);
    r.exec = nullptr;
}

void graph_clear(Engine *e) {
    for (auto &kv : e->graphs) drop(kv.second);
    e->graphs.clear();
}

// returns 1 if the pass was replayed from a graph (the caller skips its launches), 0 if the caller must issue
// them (eagerly, or into the capture this call has just opened: *captured), < 0 on error
int graph_begin(Engine *e, cudaStream_t s, const GraphKey &key, bool *captured) {
    *captured = false;
    if (!e->use_graphs || e->prof || s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread) return 0;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return 0;  // somebody else is capturing this stream: our launches simply become part of their graph
    }
    auto it = e->graphs.find(key.bytes);
    if (it == e->graphs.end()) {  // first sighting: remember it, run eagerly (one-off shapes never pay for a capture)
        if ((int)e->graphs.size() >= kMaxGraphs) {
            auto victim = e->graphs.begin();
            for (auto jt = e->graphs.begin(); jt != e->graphs.end(); ++jt)
                if (jt->second.last_use < victim->second.last_use) victim = jt;
            drop(victim->second);
            e->graphs.erase(victim);
        }
        e->graphs[key.bytes].last_use = ++e->graph_clock;
        return 0;
    }
    Engine::GraphRec &r = it->second;
    r.last_use = ++e->graph_clock;
    if (r.bad) return 0;
    if (r.exec) {
        DSRG_CUDA_TRY(cudaGraphLaunch(r.exec, s));
        e->launches += r.launches;
        e->graph_replays++;
        return 1;
    }
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        r.bad = true;
        return 0;
    }
    *captured = true;
    return 0;
}

int graph_end(Engine *e, cudaStream_t s, const GraphKey &key, bool captured, int body_rc, long long launches_before) {
    if (!captured) return body_rc;
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &g);
    Engine::GraphRec &r = e->graphs[key.bytes];
    if (body_rc != DSRG_OK || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        r.bad = true;  // never try again
        // nothing has run yet: the body's own error stands, otherwise the caller re-issues the launches eagerly
        return body_rc != DSRG_OK ? body_rc : kGraphRetry;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) {
        cudaGetLastError();
        r.bad = true;
        return kGraphRetry;
    }
    r.exec = exec;
    r.launches = e->launches - launches_before;
    DSRG_CUDA_TRY(cudaGraphLaunch(exec, s));
    e->graph_replays++;
    return DSRG_OK;
}

}  // namespace dsrg
