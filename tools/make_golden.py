"""Generate tests/golden/maze.json from the reference's own data.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tools/make_golden.py
Everything in the fixture is derived from /root/reference/computed_data/node_coords.json and the
constants written in /root/reference/pacbot_rs/src/env.rs / build.rs -- no oracle, no product
code.  These are the in-tree facts that pin behaviour."""
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "maze.json")


def main():
    nodes = [tuple(x) for x in json.load(open(os.path.join(REF, "computed_data", "node_coords.json")))]
    node_set = set(nodes)
    rows, cols = 31, 28
    # K1: wall_at(r, c) == (r, c) not in node_coords   (build.rs:78-108)
    walls = [sum(1 << c for c in range(cols) if (r, c) not in node_set) for r in range(rows)]
    # K2: VALID_ACTIONS as build.rs:141-149 computes them: [true, (x+1,y), (x-1,y), (x,y-1), (x,y+1)]
    valid_actions = [
        [True, (r + 1, c) in node_set, (r - 1, c) in node_set, (r, c - 1) in node_set, (r, c + 1) in node_set]
        for (r, c) in nodes
    ]
    degree_hist = {}
    for va in valid_actions:
        d = sum(va[1:])
        degree_hist[str(d)] = degree_hist.get(str(d), 0) + 1
    # K4: observation wall channel, obs[0, col, 30-row] = wall  (env.rs:425-426), flat [28*31]
    wall_channel = [0.0] * (28 * 31)
    for r in range(rows):
        for c in range(cols):
            wall_channel[c * 31 + (31 - r - 1)] = 0.0 if (r, c) in node_set else 1.0
    # K7: halves cleared by reset's wipe types (env.rs:190-196)
    wipe_halves = {
        "1": "col < 14", "2": "col >= 14", "3": "row < 15", "4": "row >= 15",
    }
    fixture = {
        "source": "computed_data/node_coords.json + constants of pacbot_rs/src/env.rs, build.rs",
        "rows": rows,
        "cols": cols,
        "node_coords": [list(n) for n in nodes],
        "wall_rows": walls,
        "valid_actions": valid_actions,
        "degree_histogram": degree_hist,
        "wall_channel_flat": wall_channel,
        "wall_cell_count": int(sum(wall_channel)),
        "super_pellets": [[3, 1], [3, 26], [23, 1], [23, 26]],  # env.rs:283
        "obs_shape": [17, 28, 31],  # env.rs:23
        "constants": {  # env.rs:25-38
            "TICKS_PER_UPDATE": 12, "NORMAL_TICKS_PER_STEP": 8, "MIN_TICKS_PER_STEP": 4,
            "MAX_TICKS_PER_STEP": 14, "TURN_PENALTY": -2, "LAST_REWARD": 3000,
        },
        "wipe_halves": wipe_halves,
        # Appendix G action-mask known answers at named cells
        "mask_kat": {
            "23,13": [True, False, False, True, True],
            "1,1": [True, True, False, False, True],
            "5,6": [True, True, True, True, True],
            "29,26": [True, False, True, True, False],
            "14,9": [True, True, True, True, False],
            "11,13": [True, False, False, True, True],
        },
        "four_way_nodes": [[5, 6], [5, 21], [20, 6], [20, 21]],
    }
    # self-check the hand-written KATs against the data before writing
    for key, exp in fixture["mask_kat"].items():
        r, c = map(int, key.split(","))
        assert valid_actions[nodes.index((r, c))] == exp, key
    four = sorted(list(nodes[i]) for i, va in enumerate(valid_actions) if sum(va[1:]) == 4)
    assert four == sorted(fixture["four_way_nodes"])
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(fixture, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "degree histogram", degree_hist,
          "wall cells", fixture["wall_cell_count"])


if __name__ == "__main__":
    main()
