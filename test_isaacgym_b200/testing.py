"""Test doubles for the PhysX step (NOT product code: used by ``tests/test_rollout.py`` only).

The closed-loop harness of ``rollout.py`` talks to a simulator through a three-method backend protocol
(``simulate`` / ``refresh`` / ``apply``).  Isaac Gym cannot be installed here, so the tests close the loop on these
stand-ins: ``KinematicServoBackend`` integrates the root states with the commanded velocities (what PhysX does to the
two joint-less actors of test10, whose velocity is overwritten every step), ``ReplayFrankaBackend`` holds
pre-generated gym-layout tensors.  They are not controllers: the law always runs in the CUDA library.
"""
from __future__ import annotations

import torch


class KinematicServoBackend:
    """Stand-in for ``gym.simulate`` on the test10 scene: ``pos += lin_vel * dt`` for both actors; attitude is
    whatever the controller wrote (``set_actor_root_state_tensor`` teleports the actors, test10:456)."""

    capturable = True

    def __init__(self, root_state: torch.Tensor, dt: float = 1.0 / 60.0):
        if root_state.dim() == 2:
            root_state = root_state.view(-1, 2, 13)
        self.root_state = root_state          # (N, 2, 13) fp32, actors [uav, car]  (test10:372-374)
        self.dt = float(dt)
        self.device = root_state.device

    def simulate(self) -> None:               # gym.simulate + gym.fetch_results   (test10:380-381)
        self.root_state[:, :, 0:3].add_(self.root_state[:, :, 7:10], alpha=self.dt)

    def refresh(self) -> None:                # gym.refresh_actor_root_state_tensor (test10:394)
        pass

    def apply(self) -> None:                  # gym.set_actor_root_state_tensor    (test10:456)
        pass



class ReplayFrankaBackend:
    """Stand-in for the Franka pick scene: the gym-layout tensors of ``synthetic.franka_task_inputs`` /
    ``franka_inputs`` (``jacobian (N,10,6,9)``, ``mass_matrix (N,9,9)``, ``dof_state (9N,2)``, ``rb_states (13N,13)``)
    held on the device; ``simulate`` applies the position targets kinematically to the DOF state so consecutive steps
    see different inputs.  Plumbing / aliasing test double, not a physics model."""

    capturable = True

    def __init__(self, task_inputs, franka_inputs, device):
        to = lambda t: t.to(device).clone() if isinstance(t, torch.Tensor) else t
        t = task_inputs.__class__(**{k: to(v) for k, v in task_inputs.__dict__.items()})
        f = franka_inputs.__class__(**{k: to(v) for k, v in franka_inputs.__dict__.items()})
        n = f.num_envs
        self.num_envs, self.device = n, torch.device(device)
        self.rb_states, self.dof_state = t.rb_states, t.dof_state
        self.dof_pos, self.dof_vel = t.dof_pos, t.dof_state[:, 1].view(n, 9, 1)      # franka_cube_ik_osc.py:323-326
        self.j_eef, self.mm = f.j_eef, f.mm                                          # :305-316
        self.box_idxs, self.hand_idxs = t.box_idxs, t.hand_idxs                      # :255-256,277-278
        self.init_pos, self.init_rot, self.hand_restart = t.init_pos, t.init_rot, t.hand_restart
        self.default_dof_pos = f.default_dof_pos
        self.pos_action = torch.zeros(n, 9, device=device)                           # :329
        self.effort_action = torch.zeros(n, 9, device=device)                        # :333
        self.pos_action.copy_(self.dof_pos.view(n, 9))

    def simulate(self) -> None:
        # first-order pull of the joints toward the position targets (the PhysX POS drive's visible effect)
        q = self.dof_state[:, 0].view(self.num_envs, 9)
        q.add_(self.pos_action - q, alpha=0.25)

    def refresh(self) -> None:
        pass

    def apply(self) -> None:
        pass
