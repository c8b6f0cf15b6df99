"""Device-resident policies with the reference's names and call signature
(/root/reference/src/policies.py:16-59): a policy maps (obs [B,17,28,31], action_masks [B,5])
to actions [B].  Unlike the reference nothing here touches the host: no `.cpu()`, no Python loop
over envs, no `.tolist()`."""
from __future__ import annotations

from typing import Callable

import torch

Policy = Callable[[torch.Tensor, torch.Tensor], torch.Tensor]


class MaxQPolicy:
    """argmax_a Q(s, a) over valid actions (policies.py:44-59)."""

    def __init__(self, q_net: torch.nn.Module) -> None:
        self.q_net = q_net

    @torch.no_grad()
    def q_values(self, obs: torch.Tensor) -> torch.Tensor:
        return self.q_net(obs)

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, action_masks: torch.Tensor) -> torch.Tensor:
        q = self.q_net(obs)
        return q.masked_fill(~action_masks, -torch.inf).argmax(dim=1)

    @torch.no_grad()
    def actions_u8(self, obs: torch.Tensor, action_masks: torch.Tensor) -> torch.Tensor:
        """The same choice as uint8 [B] (what the env engine steps with); a network with a fused
        head (`QNetV2.greedy_actions`) takes the masked arg-max inside its last launch."""
        greedy = getattr(self.q_net, "greedy_actions", None)
        if greedy is not None:
            return greedy(obs, action_masks)
        return self(obs, action_masks).to(torch.uint8)


class EpsilonGreedy:
    """epsilon-greedy wrapper (policies.py:16-41): greedy iff rand > epsilon, otherwise a
    uniformly random VALID action.  The wrapped policy is evaluated for the whole batch and the
    result selected with `torch.where`, so shapes are static (CUDA-graph friendly) and there is no
    host round trip."""

    def __init__(self, original_policy: Policy, num_actions: int, epsilon: float) -> None:
        self.original_policy = original_policy
        self.num_actions = num_actions
        self.epsilon = epsilon

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, action_masks: torch.Tensor) -> torch.Tensor:
        b = obs.shape[0]
        weights = action_masks.to(torch.float32)
        random_actions = torch.multinomial(weights, 1).squeeze(1)
        greedy = torch.rand(b, device=obs.device) > self.epsilon
        return torch.where(greedy, self.original_policy(obs, action_masks), random_actions)
