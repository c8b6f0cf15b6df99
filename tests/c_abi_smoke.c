/* Plain-C consumer of include/pacbot_b200.h: proves the header is valid C (no C++ / torch types)
 * and that the shared library links and answers without Python.  Built and run by
 * tests/test_c_abi.py.  Exit code 0 = ok. */
#include <stdio.h>
#include <string.h>

#include "pacbot_b200.h"

int main(void) {
    if (pb_version() != PB_VERSION) return 1;
    if (pb_sizeof_env_state() != sizeof(pb_env_state)) return 2;
    if (pb_state_bytes_per_env() != 176) return 3;
    int r = -1, c = -1;
    if (pb_maze_node(210, &r, &c) != PB_OK || r != 23 || c != 13) return 4; /* Pac-Man spawn node */
    if (pb_maze_wall_row(0) != 0xfffffffu) return 5;
    if (pb_draw_raw(1, 2, 3) != pb_draw_raw(1, 2, 3) || pb_draw_raw(1, 2, 3) == pb_draw_raw(1, 2, 4)) return 6;
    pb_engine *e = NULL;
    int rc = pb_create(&e, 64, 0, 0, 0, NULL);
    if (rc == PB_ERR_NO_DEVICE) {
        if (!strstr(pb_last_error(), "no CPU path")) return 7;
        printf("c_abi_smoke: no GPU, pb_create refused as designed\n");
        return 0;
    }
    if (rc != PB_OK) return 8;
    /* PacmanGym::new x 64 -> reset -> one step with host buffers -> obs to host */
    unsigned char actions[64], done[64], mask[64 * 5];
    int reward[64];
    static float obs[2 * PB_OBS_FLOATS];
    memset(actions, 0, sizeof actions);
    if (pb_reset(e, NULL, NULL) != PB_OK) return 9;
    if (pb_step_host(e, actions, reward, done, mask, 0, NULL, 0, NULL) != PB_OK) return 10;
    if (pb_obs_host(e, 0, 2, obs, NULL) != PB_OK) return 11;
    float walls = 0.f;
    for (int i = 0; i < 28 * 31; i++) walls += obs[i];
    if (walls != 580.0f || pb_num_envs(e) != 64) return 12;
    actions[3] = 9;
    if (pb_step_host(e, actions, reward, done, mask, 0, NULL, 0, NULL) != PB_ERR_INVALID_ACTION) return 13;
    pb_destroy(e);
    printf("c_abi_smoke: ok on a GPU\n");
    return 0;
}
