"""Maze geometry used by the host side (printing, tests, docs).

WALL_ROWS is the complement of /root/reference/computed_data/node_coords.json in the 31x28 grid
(SURVEY.md Appendix D.1); PELLET_ROWS is SURVEY.md Appendix D.2.  tests/test_tables.py checks
these literals against tests/golden/maze.json and against the tables compiled into the .so."""
from __future__ import annotations

ROWS, COLS = 31, 28

WALL_ROWS = (
    0xFFFFFFF, 0x8006001, 0xBDF6FBD, 0xBDF6FBD, 0xBDF6FBD, 0x8000001, 0xBDBFDBD, 0xBDBFDBD,
    0x8186181, 0xFDF6FBF, 0xFDF6FBF, 0xFD801BF, 0xFDBFDBF, 0xFDBFDBF, 0xFC3FC3F, 0xFDBFDBF,
    0xFDBFDBF, 0xFD801BF, 0xFDBFDBF, 0xFDBFDBF, 0x8006001, 0xBDF6FBD, 0xBDF6FBD, 0x8C00031,
    0xEDBFDB7, 0xEDBFDB7, 0x8186181, 0xBFF6FFD, 0xBFF6FFD, 0x8000001, 0xFFFFFFF,
)

PELLET_ROWS = (
    0x0000000, 0x7FF9FFE, 0x4209042, 0x4209042, 0x4209042, 0x7FFFFFE, 0x4240242, 0x4240242,
    0x7E79E7E, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x0200040,
    0x0200040, 0x0200040, 0x0200040, 0x0200040, 0x7FF9FFE, 0x4209042, 0x4209042, 0x73F9FCE,
    0x1240248, 0x1240248, 0x7E79E7E, 0x4009002, 0x4009002, 0x7FFFFFE, 0x0000000,
)

SUPER_PELLETS = ((3, 1), (3, 26), (23, 1), (23, 26))

# NODE_COORDS (pacbot_rs/src/grid.rs:13): walkable cells, row-major
NODE_COORDS = tuple(
    (r, c) for r in range(ROWS) for c in range(COLS) if not (WALL_ROWS[r] >> c) & 1
)


def wall_at(row: int, col: int) -> bool:
    if not (0 <= row < ROWS and 0 <= col < COLS):
        return True
    return bool((WALL_ROWS[row] >> col) & 1)


def action_mask_at(row: int, col: int) -> list[bool]:
    """env.rs:293-302 == build.rs:141-149: [stay, down, up, left, right]."""
    return [True, not wall_at(row + 1, col), not wall_at(row - 1, col),
            not wall_at(row, col - 1), not wall_at(row, col + 1)]
