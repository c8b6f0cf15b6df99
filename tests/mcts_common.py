"""Deterministic evaluator / root-noise used by BOTH sides of the MCTS parity tests: pure NumPy
float32 arithmetic whose every operation is exact or a single IEEE op, so the oracle (called per
leaf through ctypes) and the CUDA path (called per batch) see bit-identical numbers."""
import numpy as np


def np_evaluator(obs: np.ndarray, mask: np.ndarray):
    """(obs f32 [B,17,28,31], mask bool [B,5]) -> (value f32 [B], policy f32 [B,5])."""
    B = obs.shape[0]
    cell = obs[:, 2].reshape(B, -1).argmax(1).astype(np.int64)        # Pac-Man's cell (0 if off board)
    npel = (obs[:, 1].reshape(B, -1) != 0).sum(1).astype(np.int64)    # cells with a reward on them
    value = (cell % 13).astype(np.float32) * np.float32(3.5) - npel.astype(np.float32) * np.float32(0.25)
    w = 1 + ((cell[:, None] * (np.arange(5)[None, :] + 3)) % 11)
    w = w.astype(np.float32) * mask.astype(np.float32)
    policy = w / w.sum(1, keepdims=True)
    return value.astype(np.float32), policy.astype(np.float32)


def uniform_evaluator(obs: np.ndarray, mask: np.ndarray):
    """What train_alphazero.py's model_evaluator returns before the first update (:97-101)."""
    value = np.zeros(obs.shape[0], dtype=np.float32)
    policy = mask / mask.sum(axis=-1, keepdims=True, dtype=np.float32)
    return value, policy.astype(np.float32)


def noise_weights(tree: int, call: int, k: int) -> np.ndarray:
    """A fixed point of the simplex standing in for the Dirichlet sample of size k."""
    w = np.array([1 + ((tree * 7 + call * 13 + j * 5) % 9) for j in range(k)], np.float32)
    return w / w.sum(dtype=np.float32)


class NoiseSource:
    """Per-tree call counters; `oracle` is the callback for pymcts, `torch_fn` the hook for
    BatchedMCTS(noise_fn=...).  Both draw noise_weights(tree, call number of that tree, k)."""

    def __init__(self, n: int):
        self.calls_oracle = np.zeros(n, np.int64)
        self.calls_gpu = np.zeros(n, np.int64)

    def oracle(self, tree: int, k: int) -> np.ndarray:
        w = noise_weights(tree, int(self.calls_oracle[tree]), k)
        self.calls_oracle[tree] += 1
        return w

    def torch_fn(self, num_valid, mask):
        import torch

        nv = num_valid.cpu().numpy()
        m = np.ones(len(nv), bool) if mask is None else mask.cpu().numpy().astype(bool)
        out = np.zeros((len(nv), 5), np.float32)
        for i in np.nonzero(m)[0]:
            out[i, : nv[i]] = noise_weights(int(i), int(self.calls_gpu[i]), int(nv[i]))
            self.calls_gpu[i] += 1
        return torch.from_numpy(out).to(num_valid.device)
