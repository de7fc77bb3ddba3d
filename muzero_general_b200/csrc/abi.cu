// Host side of libmzb200.so: the C ABI declared in include/mzb200.h.
//
// Owns the device allocations (node pool, hidden-state pool, IO arenas, weight blobs), turns a
// reference state_dict into kernel layouts, and dispatches a batched MCTS.run to
//   - the fused persistent kernel (fc_search.cu) for fully-connected nets, or
//   - the step-wise pipeline select -> network -> expand+backup (tree_kernels.cu + the network
//     kernels) for residual nets, for teacher-forced tree tests and on MZ_FLAG_STEPWISE.
// No torch types, no exceptions across the boundary, never aborts.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <type_traits>
#include <string>
#include <vector>

#include "handle.h"
#include "small_search.h"

using namespace mz;

static thread_local std::string g_create_error;

int mz_fail(MzHandle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

template <typename T>
static cudaError_t dev_alloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T) + 16); }

static void* named_buffer(MzHandle* h, const char* name, size_t bytes) {
    auto it = h->named.find(name);
    if (it != h->named.end() && it->second.second >= bytes) return it->second.first;
    if (it != h->named.end()) cudaFree(it->second.first);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes + 16) != cudaSuccess) return nullptr;
    h->named[name] = {p, bytes};
    return p;
}

extern "C" int mz_abi_version(void) { return MZ_ABI_VERSION; }

extern "C" const char* mz_last_error(const MzHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int mlp_dims(const int32_t* hidden, int n_hidden, int in, int out, std::vector<int>& sizes) {
    if (n_hidden < 0 || n_hidden > MZ_MAX_LAYERS) return -1;
    sizes.clear();
    sizes.push_back(in);
    for (int i = 0; i < n_hidden; ++i) sizes.push_back(hidden[i]);
    sizes.push_back(out);
    return 0;
}

extern "C" int mz_create(const MzNetDesc* net, const MzSearchDesc* search, int device, MzHandle** out) {
    if (!net || !search || !out) return fail(nullptr, MZ_EINVAL, "mz_create: null argument");
    *out = nullptr;
    if (search->num_players > 2)       // self_play.py:429-430
        return fail(nullptr, MZ_EUNSUPPORTED, "More than two player mode not implemented.");
    if (search->num_players < 1 || search->max_games < 1 || search->num_simulations < 0)
        return fail(nullptr, MZ_EINVAL, "mz_create: bad search descriptor");
    if (net->action_space < 1 || net->action_space > MZ_MAX_ACTIONS)
        return fail(nullptr, MZ_EUNSUPPORTED, "mz_create: action_space must be in [1, 128]");
    if (net->kind != MZ_NET_FC && net->kind != MZ_NET_RESNET)
        return fail(nullptr, MZ_EUNSUPPORTED, "The network parameter should be \"fullyconnected\" or \"resnet\".");
    if (net->kind == MZ_NET_RESNET && net->downsample > 1)
        return fail(nullptr, MZ_EUNSUPPORTED, "downsample=\"CNN\" is not supported");

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= device)
        return fail(nullptr, MZ_ECUDA, "mz_create: no such CUDA device (this library has no CPU fallback)");
    MzHandle* h = new (std::nothrow) MzHandle();
    if (!h) return fail(nullptr, MZ_ENOMEM, "mz_create: out of host memory");
    h->net = *net;
    h->search = *search;
    h->search.pb_c_table = nullptr;
    h->search.sqrt_table = nullptr;
    h->search.ucb_table = nullptr;
    h->device = device;
#define MZ_CREATE_CUDA(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fail(nullptr, MZ_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));           \
            mz_destroy(h);                                                                        \
            return MZ_ECUDA;                                                                      \
        }                                                                                         \
    } while (0)
    MZ_CREATE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MZ_CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_cap = prop.sharedMemPerBlockOptin;
    MZ_CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    MZ_CREATE_CUDA(cudaEventCreate(&h->ev0));
    MZ_CREATE_CUDA(cudaEventCreate(&h->ev1));

    if (search->extra_expansions < 0) { fail(nullptr, MZ_EINVAL, "mz_create: extra_expansions < 0"); mz_destroy(h); return MZ_EINVAL; }
    // N below is the LAYOUT size of the pool and tables (room for num_simulations + 1 + extra_expansions expansions)
    const int B = search->max_games, N = search->num_simulations + search->extra_expansions, A = net->action_space;
    h->pool_n = N;
    // ---- UCB tables (self_play.py:385-390)
    std::vector<double> pbc(N + 2), sq(N + 2);
    for (int n = 0; n < N + 2; ++n) {
        pbc[n] = search->pb_c_table ? search->pb_c_table[n]
                                    : log(((double)n + search->pb_c_base + 1) / search->pb_c_base) + search->pb_c_init;
        sq[n] = search->sqrt_table ? search->sqrt_table[n] : sqrt((double)n);
    }
    MZ_CREATE_CUDA(dev_alloc(&h->d_pbc, N + 2));
    MZ_CREATE_CUDA(dev_alloc(&h->d_sqrt, N + 2));
    MZ_CREATE_CUDA(cudaMemcpy(h->d_pbc, pbc.data(), (N + 2) * 8, cudaMemcpyHostToDevice));
    MZ_CREATE_CUDA(cudaMemcpy(h->d_sqrt, sq.data(), (N + 2) * 8, cudaMemcpyHostToDevice));
    if (search->ucb_table) {
        const size_t cells = (size_t)(N + 2) * (N + 2);
        MZ_CREATE_CUDA(dev_alloc(&h->d_ucb, cells));
        MZ_CREATE_CUDA(cudaMemcpy(h->d_ucb, search->ucb_table, cells * 8, cudaMemcpyHostToDevice));
    }

    // ---- sizes
    h->obs_elems = (int64_t)net->obs_c * net->obs_h * net->obs_w;
    if (net->kind == MZ_NET_FC) {
        h->hidden_elems = net->encoding;
    } else {
        const int hh = net->downsample ? (net->obs_h + 15) / 16 : net->obs_h;
        const int hw = net->downsample ? (net->obs_w + 15) / 16 : net->obs_w;
        h->hidden_elems = (int64_t)net->channels * hh * hw;
    }
    h->pool_state_elems = h->hidden_elems;
    if (net->kind == MZ_NET_RESNET) {
        std::string e;
        h->res = resnet_create(*net, B, h->sm_count, &e);
        if (!h->res) { fail(nullptr, MZ_ECUDA, "resnet_create: " + e); mz_destroy(h); return MZ_ECUDA; }
        h->pool_state_elems = resnet_state_elems(h->res);
    }
    // ---- node pool
    const size_t slots = (size_t)B * (N + 1) * A;
    NodePool& p = h->pool;
    MZ_CREATE_CUDA(dev_alloc(&p.visit, slots));
    MZ_CREATE_CUDA(dev_alloc(&p.vsum, slots));
    MZ_CREATE_CUDA(dev_alloc(&p.mval, slots));
    MZ_CREATE_CUDA(dev_alloc(&p.reward, slots));
    MZ_CREATE_CUDA(dev_alloc(&p.prior, slots));
    MZ_CREATE_CUDA(dev_alloc(&p.expansion, slots));
    MZ_CREATE_CUDA(dev_alloc(&p.root_prior, (size_t)B * A));
    MZ_CREATE_CUDA(dev_alloc(&p.hidden, (size_t)B * (N + 1) * h->pool_state_elems));
    MZ_CREATE_CUDA(cudaMemset(p.hidden, 0, (size_t)B * (N + 1) * h->pool_state_elems * 4));   // layout padding reads as zero
    MZ_CREATE_CUDA(dev_alloc(&p.root_visit, B));
    MZ_CREATE_CUDA(dev_alloc(&p.root_vsum, B));
    MZ_CREATE_CUDA(dev_alloc(&p.root_reward, B));
    MZ_CREATE_CUDA(dev_alloc(&p.range, (size_t)B * 2));
    MZ_CREATE_CUDA(dev_alloc(&p.n_expanded, B));
    MZ_CREATE_CUDA(dev_alloc(&p.ties, B));
    MZ_CREATE_CUDA(dev_alloc(&p.max_depth, B));
    MZ_CREATE_CUDA(dev_alloc(&p.legal, (size_t)B * (MZ_MAX_ACTIONS / 32)));     // one word per game (|A| <= 32) or four (tree_wide.cu)
    MZ_CREATE_CUDA(dev_alloc(&p.path, (size_t)B * (N + 2)));
    MZ_CREATE_CUDA(dev_alloc(&p.path_reward, (size_t)B * (N + 2)));
    MZ_CREATE_CUDA(dev_alloc(&p.leaf_depth, B));
    MZ_CREATE_CUDA(dev_alloc(&p.leaf_parent, B));
    MZ_CREATE_CUDA(dev_alloc(&p.leaf_action, B));
    MZ_CREATE_CUDA(dev_alloc(&p.leaf_slot, B));
    MZ_CREATE_CUDA(dev_alloc(&p.net_value, B));
    MZ_CREATE_CUDA(dev_alloc(&p.net_reward, B));
    MZ_CREATE_CUDA(dev_alloc(&p.net_policy, (size_t)B * A));
    MZ_CREATE_CUDA(cudaMemset(p.n_expanded, 0, B * sizeof(int)));

    // ---- IO arenas (sized for max_games)
    h->in_cap = (size_t)B * (h->obs_elems * 4 + A * 1 + 4 + A * 8 + 4 + 8 + 4) + 16 * 256;
    h->out_cap = (size_t)B * (A * 4 + 8 + 4 + 4 + 4 + A * 8 + 16) + 16 * 256;
    MZ_CREATE_CUDA(cudaMalloc(&h->d_in, h->in_cap));
    MZ_CREATE_CUDA(cudaMalloc(&h->d_out, h->out_cap));
    MZ_CREATE_CUDA(cudaMallocHost(&h->h_in, h->in_cap));
    MZ_CREATE_CUDA(cudaMallocHost(&h->h_out, h->out_cap));

    const char* genv = getenv("MZ_FC_GROUP");
    if (genv) h->fc_group = atoi(genv);
    else {
        int g = 8;
        while (g < A && g < 32) g <<= 1;
        if (g < 16) g = 16;
        h->fc_group = g;                     // |A| > 32: 32 lanes stride over the outputs; the search itself is step-wise
    }
    const char* tenv = getenv("MZ_FC_THREADS");
    if (tenv) h->fc_threads = atoi(tenv);
    if (h->fc_threads < 32 || h->fc_threads > kFcMaxThreads || h->fc_threads % 32) h->fc_threads = 64;
    if ((h->fc_group < A && A <= 32) || (h->fc_group != 4 && h->fc_group != 8 && h->fc_group != 16 && h->fc_group != 32)) {
        fail(nullptr, MZ_EINVAL, "mz_create: MZ_FC_GROUP must be 4, 8, 16 or 32 and >= action_space");
        mz_destroy(h);
        return MZ_EINVAL;
    }
    *out = h;
    return MZ_OK;
}

extern "C" int mz_destroy(MzHandle* h) {
    if (!h) return MZ_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    mz_selfplay_destroy(h);
    NodePool& p = h->pool;
    void* ptrs[] = {p.visit, p.vsum, p.mval, p.reward, p.prior, p.expansion, p.root_prior, p.hidden, p.root_visit, p.root_vsum,
                    p.root_reward, p.range, p.n_expanded, p.ties, p.max_depth, p.legal, p.path, p.path_reward, p.leaf_depth,
                    p.leaf_parent, p.leaf_action, p.leaf_slot, p.net_value, p.net_reward, p.net_policy, h->d_ucb, h->d_pbc, h->d_sqrt, h->d_fc_blob, h->d_in, h->d_out};
    for (void* q : ptrs) if (q) cudaFree(q);
    for (auto& kv : h->named) cudaFree(kv.second.first);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_out) cudaFreeHost(h->h_out);
    mz_drop_graphs(h);
    if (h->res) resnet_destroy(h->res);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->part_fork) cudaEventDestroy(h->part_fork);
    for (int p = 0; p < MzHandle::kMaxParts; ++p) {
        if (h->part_join[p]) cudaEventDestroy(h->part_join[p]);
        if (h->part_stream[p]) cudaStreamDestroy(h->part_stream[p]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MZ_OK;
}

extern "C" const char* mz_numerics(const MzHandle* h) {
    if (!h) return "";
    if (h->net.kind == MZ_NET_FC) return "f32 nets + f64 tree statistics";
    return resnet_numerics(h->res);
}
extern "C" int64_t mz_hidden_elems(const MzHandle* h) { return h ? h->hidden_elems : 0; }
extern "C" int64_t mz_obs_elems(const MzHandle* h) { return h ? h->obs_elems : 0; }
extern "C" int64_t mz_launch_count(const MzHandle* h) { return h ? h->launches : 0; }
extern "C" double mz_last_search_ms(const MzHandle* h) { return h ? h->last_ms : 0.0; }
extern "C" int32_t mz_graph_partitions(const MzHandle* h) { return h ? h->graph_parts : 1; }

// ------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------
static const MzTensor* find_tensor(const MzTensor* t, int n, const std::string& name) {
    for (int i = 0; i < n; ++i) if (t[i].name && name == t[i].name) return &t[i];
    return nullptr;
}

// Appends one mlp (models.py:630-642) to the blob: per Linear the weights packed [i/4][out][i%4]
// (zero padded), the bias, and - for the first dynamics layer - the one-hot rows [A][out].
static int pack_mlp(MzHandle* h, const MzTensor* t, int n, const std::string& prefix, const std::vector<int>& sizes,
                    MlpDesc& d, std::vector<float>& blob, int onehot_rows) {
    d.n = (int)sizes.size() - 1;
    for (int l = 0; l < d.n; ++l) {
        const int in = sizes[l], out = sizes[l + 1];
        const MzTensor* w = find_tensor(t, n, prefix + "." + std::to_string(2 * l) + ".weight");
        const MzTensor* b = find_tensor(t, n, prefix + "." + std::to_string(2 * l) + ".bias");
        if (!w || !b) return fail(h, MZ_EINVAL, "mz_load_weights: missing tensor " + prefix + "." + std::to_string(2 * l));
        if (w->numel != (int64_t)in * out || b->numel != out)
            return fail(h, MZ_EINVAL, "mz_load_weights: shape mismatch for " + prefix + "." + std::to_string(2 * l));
        const int extra = (l == 0) ? onehot_rows : 0;
        const int dense = in - extra, in4 = (dense + 3) / 4;
        d.in[l] = in; d.out[l] = out; d.in_dense[l] = dense;
        while (blob.size() % 4) blob.push_back(0.0f);
        d.w_off[l] = (int)blob.size();
        blob.resize(blob.size() + (size_t)in4 * out * 4, 0.0f);
        float* dst = blob.data() + d.w_off[l];
        for (int o = 0; o < out; ++o)
            for (int i = 0; i < dense; ++i)
                dst[((size_t)(i / 4) * out + o) * 4 + (i % 4)] = w->data[(size_t)o * in + i];     // torch Linear: [out][in]
        d.b_off[l] = (int)blob.size();
        blob.insert(blob.end(), b->data, b->data + out);
        d.x_off[l] = -1;
        if (extra > 0) {
            d.x_off[l] = (int)blob.size();
            for (int a = 0; a < extra; ++a)
                for (int o = 0; o < out; ++o) blob.push_back(w->data[(size_t)o * in + dense + a]);
        }
    }
    return MZ_OK;
}

static int load_fc_weights(MzHandle* h, const MzTensor* t, int n) {
    const MzNetDesc& nd = h->net;
    const int E = nd.encoding, A = nd.action_space, F = 2 * nd.support_size + 1;
    FcNet fc{};
    std::vector<float> blob;
    std::vector<int> sz;
    int rc;
    if (mlp_dims(nd.fc_representation, nd.n_fc_representation, (int)h->obs_elems, E, sz)) return fail(h, MZ_EINVAL, "bad layers");
    if ((rc = pack_mlp(h, t, n, "representation_network.module", sz, fc.rep, blob, 0))) return rc;
    if (mlp_dims(nd.fc_dynamics, nd.n_fc_dynamics, E + A, E, sz)) return fail(h, MZ_EINVAL, "bad layers");
    if ((rc = pack_mlp(h, t, n, "dynamics_encoded_state_network.module", sz, fc.dyn, blob, A))) return rc;
    if (mlp_dims(nd.fc_reward, nd.n_fc_reward, E, F, sz)) return fail(h, MZ_EINVAL, "bad layers");
    if ((rc = pack_mlp(h, t, n, "dynamics_reward_network.module", sz, fc.rew, blob, 0))) return rc;
    if (mlp_dims(nd.fc_value, nd.n_fc_value, E, F, sz)) return fail(h, MZ_EINVAL, "bad layers");
    if ((rc = pack_mlp(h, t, n, "prediction_value_network.module", sz, fc.val, blob, 0))) return rc;
    if (mlp_dims(nd.fc_policy, nd.n_fc_policy, E, A, sz)) return fail(h, MZ_EINVAL, "bad layers");
    if ((rc = pack_mlp(h, t, n, "prediction_policy_network.module", sz, fc.pol, blob, 0))) return rc;
    fc.blob_floats = (int)blob.size();
    fc.obs_elems = (int)h->obs_elems; fc.E = E; fc.A = A; fc.S = nd.support_size; fc.F = F;
    int maxw = E > F ? E : F;
    if (A > maxw) maxw = A;
    if ((int)h->obs_elems > maxw) maxw = (int)h->obs_elems;
    while (blob.size() % 4) blob.push_back(0.0f);
    const MlpDesc* all[] = {&fc.rep, &fc.dyn, &fc.rew, &fc.val, &fc.pol};
    for (const MlpDesc* d : all) for (int l = 0; l < d->n; ++l) if (d->out[l] > maxw) maxw = d->out[l];
    fc.maxw = (maxw + 3) & ~3;
    if (h->d_fc_blob) cudaFree(h->d_fc_blob);
    h->d_fc_blob = nullptr;
    MZ_CUDA(h, dev_alloc(&h->d_fc_blob, blob.size()));
    MZ_CUDA(h, cudaMemcpy(h->d_fc_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
    h->fc = fc;
    return MZ_OK;
}

extern "C" int mz_load_weights(MzHandle* h, const MzTensor* tensors, int32_t n) {
    if (!h || !tensors || n <= 0) return fail(h, MZ_EINVAL, "mz_load_weights: null argument");
    MZ_CUDA(h, cudaSetDevice(h->device));
    MZ_CUDA(h, cudaStreamSynchronize(h->stream));
    mz_drop_graphs(h);                 // weight buffers move
    int rc;
    if (h->net.kind == MZ_NET_FC) {
        rc = load_fc_weights(h, tensors, n);
    } else {
        std::string e;
        rc = resnet_load_weights(h->res, tensors, n, &e);
        if (rc) return fail(h, rc, "mz_load_weights: " + e);
    }
    if (rc == MZ_OK) h->weights_loaded = true;
    return rc;
}

// ------------------------------------------------------------------------------------------
// search
// ------------------------------------------------------------------------------------------
struct Arena {
    size_t off = 0;
    size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; }
};

template <typename T>
static const T* stage_in(MzHandle* h, Arena& ar, const T* src, size_t count) {
    if (!src) return nullptr;
    const size_t o = ar.take(count * sizeof(T));
    memcpy(h->h_in + o, src, count * sizeof(T));
    return reinterpret_cast<const T*>(h->d_in + o);
}

struct OutSlot { void* user; size_t off, bytes; };
template <typename T>
static T* stage_out(MzHandle* h, Arena& ar, T* user, size_t count, std::vector<OutSlot>& slots) {
    if (!user) return nullptr;
    const size_t o = ar.take(count * sizeof(T));
    slots.push_back({user, o, count * sizeof(T)});
    return reinterpret_cast<T*>(h->d_out + o);
}

template <typename T>
static int debug_in(MzHandle* h, const char* name, const T* src, size_t count, int mem, const T** out) {
    *out = nullptr;
    if (!src) return MZ_OK;
    if (mem == MZ_MEM_DEVICE) { *out = src; return MZ_OK; }
    void* d = named_buffer(h, name, count * sizeof(T));
    if (!d) return fail(h, MZ_ENOMEM, std::string("out of device memory for ") + name);
    MZ_CUDA(h, cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    *out = reinterpret_cast<const T*>(d);
    return MZ_OK;
}

struct DebugOut { void* user; void* dev; size_t bytes; };
template <typename T>
static int debug_out(MzHandle* h, const char* name, T* user, size_t count, int mem, T** out, std::vector<DebugOut>& outs) {
    *out = nullptr;
    if (!user) return MZ_OK;
    if (mem == MZ_MEM_DEVICE) { *out = user; return MZ_OK; }
    void* d = named_buffer(h, name, count * sizeof(T));
    if (!d) return fail(h, MZ_ENOMEM, std::string("out of device memory for ") + name);
    outs.push_back({user, d, count * sizeof(T)});
    *out = reinterpret_cast<T*>(d);
    return MZ_OK;
}

// The x3 towers' range guard fired: fp32 CUDA-core towers from now on (the hidden-state pool is large enough for the
// dense layout; a captured graph refers to the old kernels and is dropped).
void mz_switch_to_strict(MzHandle* h) {
    if (!h->res) return;
    resnet_use_strict(h->res);
    h->pool_state_elems = resnet_state_elems(h->res);
    mz_drop_graphs(h);
    h->range_fallbacks += 1;
}

void mz_drop_graphs(MzHandle* h) {
    for (auto& e : h->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    h->graphs.clear();
}

// Number of parallel branches the captured graph of a search over n games is split into (1 = no split).
// MZ_PARTS = 1 switches the partitioned replay off, 2..4 forces that many branches wherever the kernels allow it.
static int search_partitions(MzHandle* h, int n, int continue_from) {
    const char* penv = getenv("MZ_PARTS");
    const int forced = penv ? atoi(penv) : 0;
    if (forced == 1 || continue_from > 0 || h->net.kind != MZ_NET_RESNET || h->net.action_space > 32 || !h->res) return 1;
    if (!resnet_can_partition(h->res)) return 1;
    // default: two branches, four from 1024 games on (Connect4 1024 games, N = 200: 43.3 / 39.4 / 38.8 ms per search with 1 / 2 / 4)
    int parts = forced > 1 ? std::min(forced, (int)MzHandle::kMaxParts) : (n >= 1024 ? 4 : 2);
    if (forced <= 1 && !resnet_uses_tensor_cores(h->res)) return 1;          // default: the tensor-core towers only
    while (parts > 1 && n < parts * 64) --parts;
    if (parts < 2) return 1;
    for (int p = 1; p < parts; ++p) {
        if (!h->part_stream[p] && cudaStreamCreateWithFlags(&h->part_stream[p], cudaStreamNonBlocking) != cudaSuccess) return 1;
        if (!h->part_join[p] && cudaEventCreateWithFlags(&h->part_join[p], cudaEventDisableTiming) != cudaSuccess) return 1;
    }
    if (!h->part_fork && cudaEventCreateWithFlags(&h->part_fork, cudaEventDisableTiming) != cudaSuccess) return 1;
    return parts;
}

// ------------------------------------------------------------------------------------------
// One batched search on device buffers (no synchronisation): fused FC kernel or step-wise pipeline.
// ------------------------------------------------------------------------------------------
int mz_dispatch_search(MzHandle* h, const SearchCall& call, bool teacher, bool trace, int flags) {
    const int n = call.n, N = h->search.num_simulations, A = h->net.action_space;
    int rc;
    if (h->search.extra_expansions > 0 || A > 32) flags |= MZ_FLAG_STEPWISE;     // the fused FC kernel: N + 1 expansions, one lane per action
    SearchCall cont = call;
    cont.continue_from = (flags & MZ_FLAG_CONTINUE) ? h->imported_expansions : 0;
    const SearchCall& call_ = cont;
    const bool fused = (h->net.kind == MZ_NET_FC || teacher) && !(flags & MZ_FLAG_STEPWISE);
    if (fused) {
        FcSearchArgs a{};
        a.n_games = n; a.N = N; a.A = A; a.P = h->search.num_players; a.threads = h->fc_threads;
        a.discount = h->search.discount; a.noise_frac = h->search.root_exploration_fraction; a.noise_alpha = h->search.root_dirichlet_alpha; a.seed = h->search.seed;
        a.pbc = h->d_pbc; a.sqrtn = h->d_sqrt; a.ucb = h->d_ucb;
        a.net = h->fc; a.blob = h->d_fc_blob;
        if (teacher) { a.net.E = 1; a.net.maxw = 4; a.net.blob_floats = 0; a.net.A = A; }
        a.obs = call.obs; a.legal_mask = call.legal_mask; a.to_play = call.to_play; a.add_noise = call.add_noise;
        a.noise = call.noise; a.first_index = call.first_index; a.game_id = call.game_id; a.move_index = call.move_index;
        a.visit_counts = call.visit_counts; a.root_value = call.root_value; a.root_predicted_value = call.root_predicted_value;
        a.max_tree_depth = call.max_tree_depth; a.tie_count = call.tie_count; a.root_priors = call.root_priors;
        a.value_range = call.value_range; a.teacher = call.teacher; a.trace = call.trace;
        if (call.keep_tree) a.pool = h->pool;
        FcLaunchInfo info{};
        cudaError_t e = launch_fc_search(a, h->fc_group, teacher, h->sm_count, h->smem_cap, h->stream, &info);
        if (e == cudaErrorInvalidConfiguration) {
            // the tree does not fit in shared memory next to the weights: use the HBM node pool
            (void)cudaGetLastError();
            rc = run_stepwise_search(h->net, h->search, h->pool_n, h->pool, h->d_pbc, h->d_sqrt, h->d_ucb, h->fc, h->d_fc_blob, h->res, call_,
                                     h->fc_group, h->sm_count, h->stream, &h->launches, &h->err);
            if (rc) return rc;
        } else if (e != cudaSuccess) {
            return fail(h, MZ_ECUDA, std::string("fc_search launch: ") + cudaGetErrorString(e));
        } else {
            h->launches += 1;
        }
    } else {
        // The step-wise pipeline is 16 small launches per simulation: replay it as a CUDA graph once the
        // same argument set has been seen twice (first call runs eagerly so lazy attribute setup and
        // allocations happen outside capture).  Debug modes (teacher / trace) always run eagerly.
        const bool graphable = !teacher && !trace && !kt_enabled() && getenv("MZ_NO_GRAPH") == nullptr;
        // Result pointers must not be part of a graph's identity: a caller that allocates its result arrays per call (the
        // Python engine with device memory does) would never hit the cache.  The pipeline writes into the handle's own
        // output arena and a few small device-to-device copies behind the graph hand the results over.
        struct ResultCopy { void* dst; const void* src; size_t bytes; };
        ResultCopy copies[8];
        int n_copies = 0;
        if (graphable) {
            Arena ar;
            const unsigned char* lo = h->d_out;
            const unsigned char* hi = h->d_out + h->out_cap;
            auto redirect = [&](auto*& ptr, size_t count) {
                using T = std::remove_reference_t<decltype(*ptr)>;
                const unsigned char* p8 = reinterpret_cast<const unsigned char*>(ptr);
                if (!ptr || (p8 >= lo && p8 < hi)) return;               // absent, or already in the arena (host-memory calls)
                const size_t o = ar.take(count * sizeof(T));
                if (ar.off > h->out_cap) { ar.off = o; return; }         // (cannot happen: the arena is sized for max_games)
                copies[n_copies++] = {ptr, h->d_out + o, count * sizeof(T)};
                ptr = reinterpret_cast<T*>(h->d_out + o);
            };
            redirect(cont.visit_counts, (size_t)n * A);
            redirect(cont.root_value, (size_t)n);
            redirect(cont.root_predicted_value, (size_t)n);
            redirect(cont.max_tree_depth, (size_t)n);
            redirect(cont.tie_count, (size_t)n);
            redirect(cont.root_priors, (size_t)n * A);
            redirect(cont.value_range, (size_t)n * 2);
        }
        uint64_t key = 1469598103934665603ull;
        auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
        const void* ptrs[] = {call_.obs, call_.legal_mask, call_.to_play, call_.noise, call_.first_index, call_.game_id,
                              call_.move_index, call_.visit_counts, call_.root_value, call_.root_predicted_value,
                              call_.max_tree_depth, call_.tie_count, call_.root_priors, call_.value_range};
        for (const void* q : ptrs) mix((uint64_t)(uintptr_t)q);
        mix((uint64_t)n); mix((uint64_t)call.add_noise); mix((uint64_t)call.keep_tree); mix((uint64_t)call_.continue_from);
        auto run = [&](const SearchCall& sc, cudaStream_t st) {
            return run_stepwise_search(h->net, h->search, h->pool_n, h->pool, h->d_pbc, h->d_sqrt, h->d_ucb, h->fc, h->d_fc_blob, h->res, sc,
                                       h->fc_group, h->sm_count, st, &h->launches, &h->err);
        };
        auto eager = [&]() { return run(call_, h->stream); };
        // Partitioned replay.  A simulation is a chain of dependent kernels (tower -> heads -> tower -> heads -> tree step) and
        // the tensor-core towers leave the SMs they do not fill - and every SM during the heads / tree kernels - idle.  Games
        // are independent, so the captured graph runs the simulations of P disjoint game ranges as P parallel branches: the
        // towers of one range overlap the heads and tree steps of the others.  Every array stays addressed by the global
        // game index, so the arithmetic per game - and every result - is exactly that of the whole-batch call.
        const int parts = search_partitions(h, n, call_.continue_from);
        auto partitioned = [&]() -> int {
            SearchCall root = call_;
            root.phases = kPhaseRoot;
            int rc2 = run(root, h->stream);
            if (rc2) return rc2;
            if (cudaEventRecord(h->part_fork, h->stream) != cudaSuccess) return fail(h, MZ_ECUDA, "partitioned replay: fork");
            const int per = ((n + parts - 1) / parts + 7) & ~7;
            for (int p = 0; p < parts; ++p) {
                SearchCall sc = call_;
                sc.phases = kPhaseSims;
                sc.g0 = p * per;
                sc.n = std::min(per, n - sc.g0);
                if (sc.n <= 0) continue;
                cudaStream_t st = p == 0 ? h->stream : h->part_stream[p];
                if (p > 0 && cudaStreamWaitEvent(st, h->part_fork, 0) != cudaSuccess) return fail(h, MZ_ECUDA, "partitioned replay: fork wait");
                if ((rc2 = run(sc, st))) return rc2;
                if (p > 0) {
                    if (cudaEventRecord(h->part_join[p], st) != cudaSuccess || cudaStreamWaitEvent(h->stream, h->part_join[p], 0) != cudaSuccess)
                        return fail(h, MZ_ECUDA, "partitioned replay: join");
                }
            }
            return MZ_OK;
        };
        MzHandle::SearchGraph* gr = nullptr;
        if (graphable) {
            h->graph_tick += 1;
            for (auto& e : h->graphs) if (e.key == key) gr = &e;
            if (!gr) {
                if ((int)h->graphs.size() < MzHandle::kMaxGraphs) {
                    h->graphs.emplace_back();
                    gr = &h->graphs.back();
                } else {
                    gr = &h->graphs[0];
                    for (auto& e : h->graphs) if (e.used < gr->used) gr = &e;
                    if (gr->exec) cudaGraphExecDestroy(gr->exec);
                    *gr = MzHandle::SearchGraph{};
                }
                gr->key = key;
            }
            gr->used = h->graph_tick;
        }
        if (gr && gr->exec) {
            MZ_CUDA(h, cudaGraphLaunch(gr->exec, h->stream));
            h->launches += gr->launches;
            h->graph_parts = gr->parts;
        } else if (gr && gr->seen >= 1) {
            const int64_t l0 = h->launches;
            MZ_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            rc = parts > 1 ? partitioned() : eager();
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return fail(h, MZ_ECUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&gr->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { gr->exec = nullptr; return fail(h, MZ_ECUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce)); }
            gr->launches = h->launches - l0;
            gr->parts = parts;
            h->graph_parts = parts;
            MZ_CUDA(h, cudaGraphLaunch(gr->exec, h->stream));
        } else {
            rc = eager();
            if (rc) return rc;
            if (gr) gr->seen += 1;
            h->graph_parts = 1;
        }
        for (int i = 0; i < n_copies; ++i)
            MZ_CUDA(h, cudaMemcpyAsync(copies[i].dst, copies[i].src, copies[i].bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    return MZ_OK;
}

extern "C" int mz_search(MzHandle* h, const MzSearchIO* io) {
    if (!h || !io) return fail(h, MZ_EINVAL, "mz_search: null argument");
    const int n = io->n_games, N = h->search.num_simulations, A = h->net.action_space;
    if (n < 1 || n > h->search.max_games) return fail(h, MZ_EINVAL, "mz_search: n_games out of range");
    const bool teacher = io->teacher != nullptr;
    if (!teacher && !h->weights_loaded) return fail(h, MZ_ESTATE, "mz_search: weights not loaded");
    const bool cont = (io->flags & MZ_FLAG_CONTINUE) != 0;
    if (!teacher && !io->obs && !cont) return fail(h, MZ_EINVAL, "mz_search: obs is null");
    if (cont) {
        if (n != 1) return fail(h, MZ_EINVAL, "mz_search: MZ_FLAG_CONTINUE takes one game");
        if (h->imported_expansions < 1) return fail(h, MZ_ESTATE, "mz_search: MZ_FLAG_CONTINUE without mz_import_tree");
        if (h->imported_expansions + h->search.num_simulations > h->pool_n + 1)
            return fail(h, MZ_EINVAL, "mz_search: the imported tree plus num_simulations exceeds the pool (raise extra_expansions)");
    }
    MZ_CUDA(h, cudaSetDevice(h->device));

    SearchCall call{};
    call.n = n;
    std::vector<OutSlot> outs;
    std::vector<DebugOut> dbg_outs;
    const bool host = io->mem == MZ_MEM_HOST;
    if (host) {
        Arena ai, ao;
        call.obs = (teacher || cont) ? nullptr : stage_in(h, ai, io->obs, (size_t)n * h->obs_elems);
        call.legal_mask = stage_in(h, ai, io->legal_mask, (size_t)n * A);
        call.to_play = stage_in(h, ai, io->to_play, n);
        call.noise = stage_in(h, ai, io->add_exploration_noise ? io->noise : nullptr, (size_t)n * A);
        call.first_index = stage_in(h, ai, io->first_index, n);
        call.game_id = stage_in(h, ai, io->game_id, n);
        call.move_index = stage_in(h, ai, io->move_index, n);
        if (ai.off > h->in_cap) return fail(h, MZ_EINVAL, "mz_search: input arena overflow");
        if (ai.off) MZ_CUDA(h, cudaMemcpyAsync(h->d_in, h->h_in, ai.off, cudaMemcpyHostToDevice, h->stream));
        call.visit_counts = stage_out(h, ao, io->visit_counts, (size_t)n * A, outs);
        call.root_value = stage_out(h, ao, io->root_value, n, outs);
        call.root_predicted_value = stage_out(h, ao, io->root_predicted_value, n, outs);
        call.max_tree_depth = stage_out(h, ao, io->max_tree_depth, n, outs);
        call.tie_count = stage_out(h, ao, io->tie_count, n, outs);
        call.root_priors = stage_out(h, ao, io->root_priors, (size_t)n * A, outs);
        call.value_range = stage_out(h, ao, io->value_range, (size_t)n * 2, outs);
        if (ao.off > h->out_cap) return fail(h, MZ_EINVAL, "mz_search: output arena overflow");
        call.out_bytes = ao.off;
    } else {
        call.obs = io->obs; call.legal_mask = io->legal_mask; call.to_play = io->to_play;
        call.noise = io->add_exploration_noise ? io->noise : nullptr;
        call.first_index = io->first_index; call.game_id = io->game_id; call.move_index = io->move_index;
        call.visit_counts = io->visit_counts; call.root_value = io->root_value;
        call.root_predicted_value = io->root_predicted_value; call.max_tree_depth = io->max_tree_depth;
        call.tie_count = io->tie_count; call.root_priors = io->root_priors; call.value_range = io->value_range;
    }
    call.add_noise = io->add_exploration_noise;
    int rc;
    if (teacher) {
        const MzTeacher& t = *io->teacher;
        if ((rc = debug_in(h, "t.root_value", t.root_value, n, io->mem, &call.teacher.root_value))) return rc;
        if ((rc = debug_in(h, "t.root_reward", t.root_reward, n, io->mem, &call.teacher.root_reward))) return rc;
        if ((rc = debug_in(h, "t.root_priors", t.root_priors, (size_t)n * A, io->mem, &call.teacher.root_priors))) return rc;
        if ((rc = debug_in(h, "t.value", t.value, (size_t)n * N, io->mem, &call.teacher.value))) return rc;
        if ((rc = debug_in(h, "t.reward", t.reward, (size_t)n * N, io->mem, &call.teacher.reward))) return rc;
        if ((rc = debug_in(h, "t.priors", t.priors, (size_t)n * N * A, io->mem, &call.teacher.priors))) return rc;
        if (!call.teacher.root_value || !call.teacher.root_reward || !call.teacher.root_priors ||
            (N > 0 && (!call.teacher.value || !call.teacher.reward || !call.teacher.priors)))
            return fail(h, MZ_EINVAL, "mz_search: incomplete teacher table");
    }
    if (io->trace) {
        const MzTrace& t = *io->trace;
        const int D = t.max_depth;
        call.trace.max_depth = D;
        if ((rc = debug_out(h, "r.depth", t.depth, (size_t)n * N, io->mem, &call.trace.depth, dbg_outs))) return rc;
        if ((rc = debug_out(h, "r.actions", t.actions, (size_t)n * N * D, io->mem, &call.trace.actions, dbg_outs))) return rc;
        if ((rc = debug_out(h, "r.value", t.value, (size_t)n * N, io->mem, &call.trace.value, dbg_outs))) return rc;
        if ((rc = debug_out(h, "r.reward", t.reward, (size_t)n * N, io->mem, &call.trace.reward, dbg_outs))) return rc;
        if ((rc = debug_out(h, "r.priors", t.priors, (size_t)n * N * A, io->mem, &call.trace.priors, dbg_outs))) return rc;
        if ((rc = debug_out(h, "r.root_priors_raw", t.root_priors_raw, (size_t)n * A, io->mem, &call.trace.root_priors_raw, dbg_outs))) return rc;
        if ((rc = debug_out(h, "r.root_reward", t.root_reward, n, io->mem, &call.trace.root_reward, dbg_outs))) return rc;
        if ((rc = debug_out(h, "r.noise", t.noise, (size_t)n * A, io->mem, &call.trace.noise, dbg_outs))) return rc;
        if (call.trace.depth && (!call.trace.actions || !call.trace.value || !call.trace.reward || !call.trace.priors))
            return fail(h, MZ_EINVAL, "mz_search: trace needs depth, actions, value, reward and priors together");
    }
    call.keep_tree = (io->flags & MZ_FLAG_KEEP_TREE) != 0;

    MZ_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    if ((rc = mz_dispatch_search(h, call, teacher, io->trace != nullptr, io->flags))) return rc;
    MZ_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    if (host && call.out_bytes)
        MZ_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, call.out_bytes, cudaMemcpyDeviceToHost, h->stream));
    for (const DebugOut& d : dbg_outs)
        MZ_CUDA(h, cudaMemcpyAsync(d.user, d.dev, d.bytes, cudaMemcpyDeviceToHost, h->stream));
    MZ_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->res && !teacher && resnet_take_saturations(h->res, h->stream) > 0) {
        // an activation left the fp16 range inside the tensor-core towers: the results above are outside the accuracy
        // contract.  Switch this handle to the fp32 CUDA-core towers for good and redo the search.
        mz_switch_to_strict(h);
        MZ_CUDA(h, cudaEventRecord(h->ev0, h->stream));
        if ((rc = mz_dispatch_search(h, call, teacher, io->trace != nullptr, io->flags))) return rc;
        MZ_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        if (host && call.out_bytes)
            MZ_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, call.out_bytes, cudaMemcpyDeviceToHost, h->stream));
        for (const DebugOut& d : dbg_outs)
            MZ_CUDA(h, cudaMemcpyAsync(d.user, d.dev, d.bytes, cudaMemcpyDeviceToHost, h->stream));
        MZ_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    for (const OutSlot& s : outs) memcpy(s.user, h->h_out + s.off, s.bytes);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_ms = ms;
    return MZ_OK;
}

// ------------------------------------------------------------------------------------------
// network entry points
// ------------------------------------------------------------------------------------------
static int run_inference(MzHandle* h, int n, int mem, const float* in, const int32_t* action, const MzInferenceOut* out,
                         bool recurrent) {
    if (!h || !in || !out) return fail(h, MZ_EINVAL, "inference: null argument");
    if (!h->weights_loaded) return fail(h, MZ_ESTATE, "inference: weights not loaded");
    if (n < 1) return fail(h, MZ_EINVAL, "inference: n < 1");
    if (recurrent && !action) return fail(h, MZ_EINVAL, "recurrent_inference: action is null");
    MZ_CUDA(h, cudaSetDevice(h->device));
    const int A = h->net.action_space, F = 2 * h->net.support_size + 1;
    const size_t in_elems = recurrent ? (size_t)h->hidden_elems : (size_t)h->obs_elems;
    std::vector<DebugOut> outs;
    InferCall c{};
    c.n = n; c.recurrent = recurrent;
    int rc;
    if ((rc = debug_in(h, "i.in", in, (size_t)n * in_elems, mem, &c.in))) return rc;
    if ((rc = debug_in(h, "i.action", action, n, mem, &c.action))) return rc;
    if ((rc = debug_out(h, "i.vl", out->value_logits, (size_t)n * F, mem, &c.value_logits, outs))) return rc;
    if ((rc = debug_out(h, "i.rl", out->reward_logits, (size_t)n * F, mem, &c.reward_logits, outs))) return rc;
    if ((rc = debug_out(h, "i.pl", out->policy_logits, (size_t)n * A, mem, &c.policy_logits, outs))) return rc;
    if ((rc = debug_out(h, "i.h", out->hidden, (size_t)n * h->hidden_elems, mem, &c.hidden, outs))) return rc;
    if ((rc = debug_out(h, "i.v", out->value, n, mem, &c.value, outs))) return rc;
    if ((rc = debug_out(h, "i.r", out->reward, n, mem, &c.reward, outs))) return rc;
    if (h->net.kind == MZ_NET_FC) {
        FcInferArgs a{};
        a.n = n; a.recurrent = recurrent; a.net = h->fc; a.blob = h->d_fc_blob; a.in = c.in; a.action = c.action;
        a.value_logits = c.value_logits; a.reward_logits = c.reward_logits; a.policy_logits = c.policy_logits;
        a.hidden = c.hidden; a.value = c.value; a.reward = c.reward;
        cudaError_t e = launch_fc_inference(a, h->fc_group, h->sm_count, h->stream);
        if (e != cudaSuccess) return fail(h, MZ_ECUDA, std::string("fc_inference launch: ") + cudaGetErrorString(e));
        h->launches += 1;
    } else {
        std::string e;
        rc = resnet_inference(h->res, c, h->stream, &h->launches, &e);
        if (rc) return fail(h, rc, "resnet_inference: " + e);
    }
    if (h->res && resnet_take_saturations(h->res, h->stream) > 0) {
        mz_switch_to_strict(h);
        std::string e;
        rc = resnet_inference(h->res, c, h->stream, &h->launches, &e);
        if (rc) return fail(h, rc, "resnet_inference: " + e);
    }
    for (const DebugOut& d : outs)
        MZ_CUDA(h, cudaMemcpyAsync(d.user, d.dev, d.bytes, cudaMemcpyDeviceToHost, h->stream));
    MZ_CUDA(h, cudaStreamSynchronize(h->stream));
    return MZ_OK;
}

extern "C" int mz_initial_inference(MzHandle* h, int32_t n, int32_t mem, const float* obs, const MzInferenceOut* out) {
    return run_inference(h, n, mem, obs, nullptr, out, false);
}

extern "C" int mz_recurrent_inference(MzHandle* h, int32_t n, int32_t mem, const float* hidden, const int32_t* action,
                                      const MzInferenceOut* out) {
    return run_inference(h, n, mem, hidden, action, out, true);
}

// ------------------------------------------------------------------------------------------
// tree export
// ------------------------------------------------------------------------------------------
extern "C" int mz_export_tree(MzHandle* h, int32_t game, MzTreeExport* out) {
    if (!h || !out) return fail(h, MZ_EINVAL, "mz_export_tree: null argument");
    if (game < 0 || game >= h->search.max_games) return fail(h, MZ_EINVAL, "mz_export_tree: game out of range");
    MZ_CUDA(h, cudaSetDevice(h->device));
    MZ_CUDA(h, cudaStreamSynchronize(h->stream));
    const int N = h->pool_n, A = h->net.action_space;
    const size_t S = (size_t)(N + 1) * A, base = (size_t)game * S;
    const NodePool& p = h->pool;
    int nexp = 0;
    MZ_CUDA(h, cudaMemcpy(&nexp, p.n_expanded + game, 4, cudaMemcpyDeviceToHost));
    if (nexp < 1) return fail(h, MZ_ESTATE, "mz_export_tree: no tree kept for this game (use MZ_FLAG_KEEP_TREE)");
    out->n_expansions = nexp;
    const size_t used = (size_t)nexp * A;
    if (out->child_visit) MZ_CUDA(h, cudaMemcpy(out->child_visit, p.visit + base, used * 4, cudaMemcpyDeviceToHost));
    if (out->child_value_sum) MZ_CUDA(h, cudaMemcpy(out->child_value_sum, p.vsum + base, used * 8, cudaMemcpyDeviceToHost));
    if (out->child_reward) MZ_CUDA(h, cudaMemcpy(out->child_reward, p.reward + base, used * 4, cudaMemcpyDeviceToHost));
    if (out->child_expansion) MZ_CUDA(h, cudaMemcpy(out->child_expansion, p.expansion + base, used * 4, cudaMemcpyDeviceToHost));
    if (out->child_prior) {
        std::vector<float> pf(used);
        std::vector<double> rp(A);
        MZ_CUDA(h, cudaMemcpy(pf.data(), p.prior + base, used * 4, cudaMemcpyDeviceToHost));
        MZ_CUDA(h, cudaMemcpy(rp.data(), p.root_prior + (size_t)game * A, A * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < used; ++i) out->child_prior[i] = (i < (size_t)A) ? rp[i] : (double)pf[i];
    }
    if (out->hidden) {
        const float* src = p.hidden + (size_t)game * (N + 1) * h->pool_state_elems;
        if (h->net.kind == MZ_NET_RESNET) {
            void* tmp = named_buffer(h, "x.hidden", (size_t)nexp * h->hidden_elems * 4);
            if (!tmp) return fail(h, MZ_ENOMEM, "mz_export_tree: out of device memory");
            if (resnet_states_to_nchw(h->res, src, nexp, (float*)tmp, h->stream)) return fail(h, MZ_ECUDA, "mz_export_tree: layout conversion failed");
            MZ_CUDA(h, cudaStreamSynchronize(h->stream));
            src = (const float*)tmp;
        }
        MZ_CUDA(h, cudaMemcpy(out->hidden, src, (size_t)nexp * h->hidden_elems * 4, cudaMemcpyDeviceToHost));
    }
    MZ_CUDA(h, cudaMemcpy(&out->root_visit, p.root_visit + game, 4, cudaMemcpyDeviceToHost));
    MZ_CUDA(h, cudaMemcpy(&out->root_value_sum, p.root_vsum + game, 8, cudaMemcpyDeviceToHost));
    MZ_CUDA(h, cudaMemcpy(&out->root_reward, p.root_reward + game, 4, cudaMemcpyDeviceToHost));
    return MZ_OK;
}

// ------------------------------------------------------------------------------------------
// tree import (override_root_with, self_play.py:275-277)
// ------------------------------------------------------------------------------------------
extern "C" int mz_import_tree(MzHandle* h, int32_t game, const MzTreeExport* t) {
    if (!h || !t) return fail(h, MZ_EINVAL, "mz_import_tree: null argument");
    if (game < 0 || game >= h->search.max_games) return fail(h, MZ_EINVAL, "mz_import_tree: game out of range");
    const int N = h->pool_n, A = h->net.action_space, K = t->n_expansions;
    if (K < 1 || K > N + 1) return fail(h, MZ_EINVAL, "mz_import_tree: n_expansions does not fit the pool (raise extra_expansions)");
    if (!t->child_visit || !t->child_value_sum || !t->child_reward || !t->child_prior || !t->child_expansion)
        return fail(h, MZ_EINVAL, "mz_import_tree: incomplete tree");
    MZ_CUDA(h, cudaSetDevice(h->device));
    MZ_CUDA(h, cudaStreamSynchronize(h->stream));
    const size_t S = (size_t)(N + 1) * A, base = (size_t)game * S, used = (size_t)K * A;
    const NodePool& p = h->pool;
    std::vector<float> prior(used);
    std::vector<double> mval(used, 0.0), rp(A);
    const double discount = h->search.discount;
    const bool two = h->search.num_players == 2;
    for (size_t i = 0; i < used; ++i) {
        prior[i] = (float)t->child_prior[i];
        if (t->child_visit[i] > 0) {
            // the value term selection reads back: reward + discount * (+/-)(value_sum / visits)   (tree.cuh, tree_backup)
            const double q = t->child_value_sum[i] / (double)t->child_visit[i];
            mval[i] = (double)t->child_reward[i] + discount * (two ? -q : q);
        }
    }
    for (int a = 0; a < A; ++a) rp[a] = t->child_prior[a];
    MZ_CUDA(h, cudaMemcpy(p.visit + base, t->child_visit, used * 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.vsum + base, t->child_value_sum, used * 8, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.mval + base, mval.data(), used * 8, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.reward + base, t->child_reward, used * 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.prior + base, prior.data(), used * 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.expansion + base, t->child_expansion, used * 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.root_prior + (size_t)game * A, rp.data(), A * 8, cudaMemcpyHostToDevice));
    // children of a non-root node: the whole action space (one mask word per game, four for |A| > 32)
    const double range[2] = {INFINITY, -INFINITY};
    const int zero = 0;
    if (A > 32) {
        unsigned words[MZ_MAX_ACTIONS / 32];
        for (int j = 0; j < MZ_MAX_ACTIONS / 32; ++j) {
            const int left = A - 32 * j;
            words[j] = left >= 32 ? 0xffffffffu : (left > 0 ? ((1u << left) - 1u) : 0u);
        }
        MZ_CUDA(h, cudaMemcpy(p.legal + (size_t)game * (MZ_MAX_ACTIONS / 32), words, sizeof(words), cudaMemcpyHostToDevice));
    } else {
        const unsigned legal = A >= 32 ? 0xffffffffu : ((1u << A) - 1u);
        MZ_CUDA(h, cudaMemcpy(p.legal + game, &legal, 4, cudaMemcpyHostToDevice));
    }
    MZ_CUDA(h, cudaMemcpy(p.root_visit + game, &t->root_visit, 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.root_vsum + game, &t->root_value_sum, 8, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.root_reward + game, &t->root_reward, 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.range + 2 * (size_t)game, range, 16, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.n_expanded + game, &K, 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.ties + game, &zero, 4, cudaMemcpyHostToDevice));
    MZ_CUDA(h, cudaMemcpy(p.max_depth + game, &zero, 4, cudaMemcpyHostToDevice));
    if (t->hidden) {
        float* dst = p.hidden + (size_t)game * (N + 1) * h->pool_state_elems;
        if (h->net.kind == MZ_NET_RESNET) {
            void* tmp = named_buffer(h, "x.import", (size_t)K * h->hidden_elems * 4);
            if (!tmp) return fail(h, MZ_ENOMEM, "mz_import_tree: out of device memory");
            MZ_CUDA(h, cudaMemcpy(tmp, t->hidden, (size_t)K * h->hidden_elems * 4, cudaMemcpyHostToDevice));
            if (resnet_states_from_nchw(h->res, (const float*)tmp, K, dst, h->stream)) return fail(h, MZ_ECUDA, "mz_import_tree: layout conversion failed");
            MZ_CUDA(h, cudaStreamSynchronize(h->stream));
        } else {
            MZ_CUDA(h, cudaMemcpy(dst, t->hidden, (size_t)K * h->hidden_elems * 4, cudaMemcpyHostToDevice));
        }
    }
    h->imported_expansions = K;
    return MZ_OK;
}

// ------------------------------------------------------------------------------------------
// per-kernel-class device times for bench.py's roofline line (ktimer.h)
extern "C" int mz_kernel_timing(MzHandle* h, int32_t enable) {
    if (!h) return fail(nullptr, MZ_EINVAL, "mz_kernel_timing: null handle");
    kt_enable(enable != 0);
    return MZ_OK;
}

extern "C" int mz_kernel_times(MzHandle* h, double* ms, int64_t* count) {
    if (!h || !ms || !count) return fail(h, MZ_EINVAL, "mz_kernel_times: bad argument");
    cudaSetDevice(h->device);
    MZ_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < MZ_KERNEL_CLASSES; ++i) { ms[i] = 0.0; count[i] = 0; }
    cudaError_t e = kt_collect(ms, count);
    if (e != cudaSuccess) return fail(h, MZ_ECUDA, std::string("mz_kernel_times: ") + cudaGetErrorString(e));
    return MZ_OK;
}

extern "C" int mz_debug_small_search_plan(int32_t H, int32_t W, int32_t C, int32_t A, int32_t n, int32_t sm_count, int32_t tower_floats,
                                          int32_t heads_floats, int32_t scratch_floats, int32_t cap_channels, int64_t* plan) {
    if (!plan) return 0;
    int P, CO, G, tile, threads, row_stride, board_stride;
    size_t smem;
    if (!small_search_shape(H, W, C, A, n, sm_count, tower_floats, heads_floats, scratch_floats, cap_channels, &P, &CO, &G, &tile, &threads,
                            &smem, &row_stride, &board_stride))
        return 0;
    const int64_t out[8] = {P, CO, G, tile, threads, (int64_t)smem, row_stride, board_stride};
    for (int i = 0; i < 8; ++i) plan[i] = out[i];
    return 1;
}

// debug: one conv3x3 through either implementation (host NCHW in / out)
// ------------------------------------------------------------------------------------------
extern "C" int mz_debug_conv3x3(int device, int32_t n, int32_t C, int32_t H, int32_t W, const float* x, const float* w,
                                const float* bias, const float* residual, int32_t relu, int32_t use_tensor_cores, float* out) {
    if (!x || !w || !out || n < 1) return fail(nullptr, MZ_EINVAL, "mz_debug_conv3x3: bad argument");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, MZ_ECUDA, "mz_debug_conv3x3: no such device");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, MZ_ECUDA, "mz_debug_conv3x3: device query failed");
    std::string e;
    int rc = resnet_debug_conv(n, C, H, W, x, w, bias, residual, relu, use_tensor_cores, out, prop.multiProcessorCount, &e);
    if (rc) return fail(nullptr, rc, "mz_debug_conv3x3: " + e);
    return MZ_OK;
}
