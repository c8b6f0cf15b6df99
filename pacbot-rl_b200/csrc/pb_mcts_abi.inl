// pb_mcts_abi.inl -- host side of the pb_mcts_* / pb_copy_envs entry points
// (include/pacbot_b200.h); included at the end of pb_kernels.cu, after `struct pb_engine`.

struct pb_mcts {
    pb_engine *root = nullptr;  // borrowed
    pb_engine *sim = nullptr;   // borrowed; slot i holds the env clone tree i simulates with
    int device = 0;
    MctsView v{};
    // pb_mcts_backprop_collect's report to the host: arrival counter and launch counter (device),
    // the published word (pinned host memory); created on first use
    unsigned int *arrive = nullptr;
    unsigned long long *launches = nullptr;
    unsigned long long *host_word = nullptr;
};

namespace {
unsigned mcts_grid(const MctsView &v) { return (unsigned)((v.n + v.tpw - 1) / v.tpw); }
}  // namespace

extern "C" {

int pb_mcts_create(pb_mcts **out, pb_engine *root, pb_engine *sim, int max_nodes) {
    if (!out || !root || !sim || root == sim || root->n != sim->n || root->device != sim->device ||
        max_nodes < 2 || max_nodes > 32767)
        return fail(PB_ERR_INVALID_ARG,
                    "pb_mcts_create: need two distinct engines of equal size on one device and "
                    "2 <= max_nodes <= 32767");
    DeviceGuard g(root->device);
    if (!g.ok) return fail(PB_ERR_NO_DEVICE, "pb_mcts_create: cannot select the CUDA device");
    pb_mcts *m = new pb_mcts();
    m->root = root;
    m->sim = sim;
    m->device = root->device;
    MctsView &v = m->v;
    v.n = root->n;
    v.L = max_nodes;
    // physical slots per tree and pool: 8x the live capacity (a compaction about every tenth
    // move instead of every move), PB_MCTS_POOL_FACTOR overrides; node indices stay below 65536
    int factor = 8;
    if (const char *f = getenv("PB_MCTS_POOL_FACTOR")) factor = atoi(f) > 0 ? atoi(f) : 1;
    long long phys = (long long)max_nodes * factor;
    v.M = (int)(phys > 65535 ? 65535 : phys);
    // trees per warp: 1 while every tree can have a warp of its own (about 12 resident warps per
    // SM at 160 registers per thread), else the next power of two; PB_MCTS_TREES_PER_WARP overrides
    const long long warp_slots = (long long)root->num_sms * 12;
    int tpw = 1;
    while (tpw < 32 && (v.n + tpw - 1) / tpw > warp_slots) tpw *= 2;
    if (const char *t = getenv("PB_MCTS_TREES_PER_WARP")) {
        const int req = atoi(t);
        if (req >= 1 && req <= 32) tpw = req;
    }
    v.tpw = tpw;
    const size_t n = (size_t)v.n, L = (size_t)max_nodes, M = (size_t)v.M;
    cudaError_t err = cudaSuccess;
    auto alloc = [&](auto **p, size_t bytes, int fill) {
        if (err != cudaSuccess) return;
        err = cudaMalloc(reinterpret_cast<void **>(p), bytes);
        if (err == cudaSuccess) err = cudaMemset(*p, fill, bytes);
    };
    alloc(&v.pool, sizeof(MctsNode) * 2 * n * M, 0);
    alloc(&v.cur, sizeof(int32_t) * n, 0);
    alloc(&v.root, sizeof(int32_t) * n, 0);
    alloc(&v.n_alloc, sizeof(int32_t) * n, 0);
    alloc(&v.traj_len, sizeof(int32_t) * n, 0);
    alloc(&v.path, sizeof(int32_t) * n * L, 0);
    alloc(&v.traj_act, n * L, 0);
    alloc(&v.traj_rew, sizeof(int32_t) * n * L, 0);
    alloc(&v.leaf_kind, n, 0);
    alloc(&v.stack, sizeof(int32_t) * n * L, 0);
    alloc(&v.epoch, sizeof(unsigned long long), 0);
    alloc(&v.noise_calls, sizeof(int32_t) * n, 0);
    alloc(&v.error, sizeof(unsigned int), 0);
    if (err != cudaSuccess) {
        std::string msg = std::string("pb_mcts_create: ") + cudaGetErrorString(err);
        pb_mcts_destroy(m);
        return fail(PB_ERR_CUDA, msg);
    }
    *out = m;
    return PB_OK;
}

int pb_mcts_destroy(pb_mcts *m) {
    if (!m) return PB_OK;
    DeviceGuard g(m->device);
    MctsView &v = m->v;
    cudaFree(v.pool);
    cudaFree(v.cur);
    cudaFree(v.root);
    cudaFree(v.n_alloc);
    cudaFree(v.traj_len);
    cudaFree(v.path);
    cudaFree(v.traj_act);
    cudaFree(v.traj_rew);
    cudaFree(v.leaf_kind);
    cudaFree(v.stack);
    cudaFree(v.epoch);
    cudaFree(v.noise_calls);
    cudaFree(v.error);
    cudaFree(m->arrive);
    cudaFree(m->launches);
    if (m->host_word) cudaFreeHost(m->host_word);
    delete m;
    return PB_OK;
}

int pb_mcts_max_nodes(const pb_mcts *m) { return m ? m->v.L : 0; }
uint64_t pb_sim_seed(uint64_t seed, uint64_t epoch) { return sim_seed(seed, epoch); }

int pb_mcts_reset_root(pb_mcts *m, const uint8_t *mask_dev, const float *value_dev,
                       const float *policy_dev, void *stream) {
    if (!m || !value_dev || !policy_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_mcts_reset_root: null argument");
    DeviceGuard g(m->device);
    k_mcts_reset_root<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(
        m->v, mask_dev, value_dev, policy_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_root_noise(pb_mcts *m, const uint8_t *mask_dev, const float *noise_dev, float mix,
                       void *stream) {
    if (!m || !noise_dev) return fail(PB_ERR_INVALID_ARG, "pb_mcts_root_noise: null argument");
    DeviceGuard g(m->device);
    k_mcts_root_noise<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(
        m->v, m->root->st, mask_dev, noise_dev, 1.0f - mix, mix);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_root_noise_dirichlet(pb_mcts *m, const uint8_t *mask_dev, float alpha, float mix, void *stream) {
    if (!m || !(alpha > 0.0f)) return fail(PB_ERR_INVALID_ARG, "pb_mcts_root_noise_dirichlet: bad arguments");
    DeviceGuard g(m->device);
    k_mcts_root_noise_dirichlet<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(
        m->v, m->root->st, mask_dev, alpha, 1.0f - mix, mix, m->root->seed, m->root->gid_base);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_select(pb_mcts *m, const uint8_t *active_dev, uint8_t *leaf_kind_dev,
                   uint8_t *leaf_mask_dev, void *stream) {
    if (!m) return fail(PB_ERR_INVALID_ARG, "pb_mcts_select: null handle");
    DeviceGuard g(m->device);
    k_mcts_select<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(
        m->v, m->root->st, m->sim->st, m->root->seed, m->root->gid_base, active_dev, leaf_kind_dev,
        leaf_mask_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_backprop(pb_mcts *m, const float *value_dev, const float *policy_dev, void *stream) {
    if (!m || !value_dev || !policy_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_mcts_backprop: null argument");
    DeviceGuard g(m->device);
    k_mcts_backprop<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(m->v, value_dev,
                                                                                  policy_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_take_action(pb_mcts *m, const uint8_t *mask_dev, const uint8_t *actions_dev,
                        int32_t *reward_dev, uint8_t *done_dev, uint8_t *status_dev, void *stream) {
    if (!m || !actions_dev || !reward_dev || !done_dev || !status_dev)
        return fail(PB_ERR_INVALID_ARG, "pb_mcts_take_action: null argument");
    DeviceGuard g(m->device);
    k_mcts_take_action<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(
        m->v, m->root->st, rng_params(m->root), mask_dev, actions_dev, reward_dev, done_dev,
        status_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_collect_act(pb_mcts *m, pb_engine *snap, int max_episode_length, int tree_size,
                        int32_t *remaining_dev, int32_t *ep_len_dev, uint8_t *ep_mask_dev,
                        float *ep_dist_dev, float *ep_value_dev, uint8_t *status_dev,
                        uint32_t *flags_dev, void *stream) {
    if (!m || !snap || !remaining_dev || !ep_len_dev || !ep_mask_dev || !ep_dist_dev ||
        !ep_value_dev || !status_dev || !flags_dev || max_episode_length < 1 || tree_size < 2 ||
        snap->device != m->device || snap->n < m->v.n * (long long)max_episode_length)
        return fail(PB_ERR_INVALID_ARG,
                    "pb_mcts_collect_act: bad arguments (snap needs n * max_episode_length envs)");
    DeviceGuard g(m->device);
    k_mcts_collect_act<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(
        m->v, m->root->st, rng_params(m->root), snap->st, max_episode_length, tree_size,
        remaining_dev, ep_len_dev, ep_mask_dev, ep_dist_dev, ep_value_dev, status_dev, flags_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_step_word(pb_mcts *m, uint64_t **host_word_out) {
    if (!m || !host_word_out) return fail(PB_ERR_INVALID_ARG, "pb_mcts_step_word: null argument");
    DeviceGuard g(m->device);
    if (!m->host_word) {
        PB_CUDA(cudaMalloc(&m->arrive, sizeof(unsigned int)));
        PB_CUDA(cudaMemset(m->arrive, 0, sizeof(unsigned int)));
        PB_CUDA(cudaMalloc(&m->launches, sizeof(unsigned long long)));
        PB_CUDA(cudaMemset(m->launches, 0, sizeof(unsigned long long)));
        PB_CUDA(cudaMallocHost(&m->host_word, sizeof(unsigned long long)));
        *m->host_word = 0ull;
    }
    *host_word_out = reinterpret_cast<uint64_t *>(m->host_word);
    return PB_OK;
}

int pb_mcts_backprop_collect(pb_mcts *m, const float *value_dev, const float *policy_dev, pb_engine *snap,
                             int max_episode_length, int tree_size, int32_t *remaining_dev,
                             int32_t *ep_len_dev, uint8_t *ep_mask_dev, float *ep_dist_dev,
                             float *ep_value_dev, uint8_t *status_dev, uint32_t *flags_dev, float alpha,
                             float mix, int publish, void *stream) {
    if (!m || !value_dev || !policy_dev || !snap || !remaining_dev || !ep_len_dev || !ep_mask_dev ||
        !ep_dist_dev || !ep_value_dev || !status_dev || !flags_dev || max_episode_length < 1 ||
        tree_size < 2 || !(alpha > 0.0f) || snap->device != m->device ||
        snap->n < m->v.n * (long long)max_episode_length)
        return fail(PB_ERR_INVALID_ARG,
                    "pb_mcts_backprop_collect: bad arguments (snap needs n * max_episode_length envs)");
    if (publish && !m->host_word)
        return fail(PB_ERR_INVALID_ARG, "pb_mcts_backprop_collect: publish needs pb_mcts_step_word first");
    DeviceGuard g(m->device);
    k_mcts_backprop_collect<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(
        m->v, value_dev, policy_dev, m->root->st, rng_params(m->root), snap->st, max_episode_length,
        tree_size, remaining_dev, ep_len_dev, ep_mask_dev, ep_dist_dev, ep_value_dev, status_dev, flags_dev,
        alpha, 1.0f - mix, mix, m->root->seed, m->root->gid_base, m->arrive, m->launches,
        publish ? m->host_word : nullptr);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_query(pb_mcts *m, int what, void *out_dev, void *stream) {
    if (!m || !out_dev || what < 0 || what > MCTS_Q_ROOT_VISITS)
        return fail(PB_ERR_INVALID_ARG, "pb_mcts_query: bad arguments");
    DeviceGuard g(m->device);
    k_mcts_query<<<mcts_grid(m->v), MCTS_THREADS, 0, (cudaStream_t)stream>>>(m->v, m->root->st,
                                                                               what, out_dev);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int pb_mcts_check(pb_mcts *m, uint64_t *epoch_out, void *stream) {
    if (!m) return fail(PB_ERR_INVALID_ARG, "pb_mcts_check: null handle");
    DeviceGuard g(m->device);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned int err = 0;
    unsigned long long epoch = 0;
    PB_CUDA(cudaMemcpyAsync(&err, m->v.error, sizeof(err), cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaMemcpyAsync(&epoch, m->v.epoch, sizeof(epoch), cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
    if (epoch_out) *epoch_out = epoch;
    if (err) {
        PB_CUDA(cudaMemsetAsync(m->v.error, 0, sizeof(err), s));
        std::string msg = "pb_mcts:";
        if (err & MCTS_ERR_POOL_FULL) msg += " node pool full (search beyond max_nodes);";
        if (err & MCTS_ERR_EXPAND_TWICE) msg += " tried to expand the same child twice;";
        if (err & MCTS_ERR_ROOT_DONE) msg += " search from a finished env;";
        return fail(PB_ERR_INVALID_ARG, msg);
    }
    return PB_OK;
}

int pb_copy_envs(pb_engine *dst, const int64_t *dst_idx_dev, const pb_engine *src,
                 const int64_t *src_idx_dev, int64_t count, void *stream) {
    if (!dst || !src || count < 0 || dst->device != src->device)
        return fail(PB_ERR_INVALID_ARG, "pb_copy_envs: bad arguments");
    if ((!dst_idx_dev && count > dst->n) || (!src_idx_dev && count > src->n))
        return fail(PB_ERR_INVALID_ARG, "pb_copy_envs: count exceeds the engine size");
    if (count == 0) return PB_OK;
    DeviceGuard g(dst->device);
    const unsigned blocks = (unsigned)((count * 32 + 127) / 128);
    k_copy_envs<<<blocks, 128, 0, (cudaStream_t)stream>>>(
        dst->st, reinterpret_cast<const long long *>(dst_idx_dev), src->st,
        reinterpret_cast<const long long *>(src_idx_dev), count);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // extern "C"
