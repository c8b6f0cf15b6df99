// host_engine.cpp -- TEST INFRASTRUCTURE: the device game logic of the product
// (pacbot-rl_b200/csrc/pb_engine.cuh: gym_step / gym_reset / action_mask_bits, the functions
// k_step and k_reset call per thread) compiled for the HOST with g++ through the shim
// cuda_runtime.h next to this file, so that tests/test_host_engine.py can run long differential
// campaigns of the *same source the GPU executes* against the scalar C oracle in a container
// without a GPU.  It mirrors what k_step / k_reset do around those calls
// (pacbot-rl_b200/csrc/pb_kernels.cu) and converts to/from pb_env_state exactly like
// pb_get_state / pb_set_state.  It is NOT a CPU path of the product: nothing under
// pacbot-rl_b200/ builds, loads or links it.
#include "host_common.h"

using namespace pb;

namespace {

using namespace host_common;

RngCtx make_rng(uint64_t seed, int64_t gid_base, const uint32_t *tape, int64_t tape_len,
                int64_t i) {
    RngCtx r;
    r.seed = seed;
    r.gid = (uint64_t)(gid_base + i);
    r.tape = tape ? tape + i * tape_len : nullptr;
    r.tape_len = (uint32_t)tape_len;
    return r;
}

void write_mask(uint8_t *out, uint32_t m) {
    for (int k = 0; k < 5; k++) out[k] = (m >> k) & 1u;
}

}  // namespace

extern "C" {

size_t he_sizeof_env_state(void) { return sizeof(pb_env_state); }

// what one k_step thread does (pb_kernels.cu k_step, actions given by the caller)
int64_t he_step(pb_env_state *st, int64_t n, const uint8_t *actions, int32_t *reward,
                uint8_t *done_out, uint8_t *mask_out, int auto_reset, uint64_t seed,
                int64_t gid_base, const uint32_t *tape, int64_t tape_len) {
    int64_t bad = -1;
    for (int64_t i = 0; i < n; i++) {
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(st[i], s, f);
        const int action = actions[i];
        if (action > 4) {
            if (bad < 0) bad = i;
            reward[i] = 0;
            done_out[i] = 0;
            if (mask_out) write_mask(mask_out + i * 5, action_mask_bits(s, H_TABLES.walls));
            continue;
        }
        const PelRef pel{f, 1};
        const RngCtx rng = make_rng(seed, gid_base, tape, tape_len, i);
        int rew;
        bool done;
        gym_step(s, pel, H_TABLES, H_TABLES.walls, rng, action, rew, done);
        if (done && auto_reset) gym_reset(s, pel, H_TABLES, H_TABLES.walls, rng, true);
        to_host(s, f, st[i]);
        reward[i] = rew;
        done_out[i] = done ? 1 : 0;
        if (mask_out) write_mask(mask_out + i * 5, action_mask_bits(s, H_TABLES.walls));
    }
    return bad;
}

// what one k_reset thread does; randomise == 0 is the constructor state (pb_create)
void he_reset(pb_env_state *st, int64_t n, const uint8_t *mask, int randomise, uint64_t seed,
              int64_t gid_base, const uint32_t *tape, int64_t tape_len) {
    for (int64_t i = 0; i < n; i++) {
        if (mask && !mask[i]) continue;
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(st[i], s, f);
        const PelRef pel{f, 1};
        gym_reset(s, pel, H_TABLES, H_TABLES.walls,
                  make_rng(seed, gid_base, tape, tape_len, i), randomise != 0);
        to_host(s, f, st[i]);
    }
}

void he_action_mask(const pb_env_state *st, int64_t n, uint8_t *out) {
    for (int64_t i = 0; i < n; i++) {
        EnvR s;
        uint32_t f[PEL_WORDS];
        from_host(st[i], s, f);
        write_mask(out + i * 5, action_mask_bits(s, H_TABLES.walls));
    }
}

}  // extern "C"
