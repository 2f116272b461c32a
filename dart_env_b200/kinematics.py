"""Batched forward kinematics in torch for the host-side task layers that read body poses
(`bodynode.to_world()`, `.com()` in the reference envs whose obs/reward are not fused in the kernel).

General 3-D (revolute / prismatic / weld), DART convention T_child = T_parent * Tpj * J(q) * Tcj^-1."""
from __future__ import annotations

import numpy as np
import torch

from .skel import JOINT_PRISMATIC, JOINT_REVOLUTE, Model, inv_transform


def _rot_axis(axis: torch.Tensor, th: torch.Tensor) -> torch.Tensor:
    """Rodrigues: rotation by th[N] about the unit vector axis[3] -> [N,3,3]."""
    a = axis.to(th.dtype)
    K = torch.tensor([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]], dtype=th.dtype, device=th.device)
    eye = torch.eye(3, dtype=th.dtype, device=th.device)
    s, c = torch.sin(th)[:, None, None], torch.cos(th)[:, None, None]
    return eye + s * K + (1 - c) * (K @ K)


def body_transforms(model: Model, q: torch.Tensor):
    """World rotation [N,nb,3,3] and origin [N,nb,3] of every DART bodynode for positions q[N,nd]."""
    n, dev, dt = q.shape[0], q.device, torch.float64
    q = q.to(dt)
    Rs, ps = [], []
    for b in model.bodies:
        Tpj = torch.tensor(b.T_parent_joint, dtype=dt, device=dev)
        Tcji = torch.tensor(inv_transform(b.T_child_joint), dtype=dt, device=dev)
        Rj = torch.eye(3, dtype=dt, device=dev).expand(n, 3, 3)
        pj = torch.zeros((n, 3), dtype=dt, device=dev)
        if b.joint_type == JOINT_REVOLUTE:
            Rj = _rot_axis(torch.tensor(b.axis, dtype=dt, device=dev), q[:, b.dof])
        elif b.joint_type == JOINT_PRISMATIC:
            pj = q[:, b.dof, None] * torch.tensor(b.axis, dtype=dt, device=dev)[None, :]
        # rel = Tpj * J * Tcj^-1
        R = Tpj[:3, :3] @ Rj @ Tcji[:3, :3]
        p = (Tpj[:3, :3] @ (Rj @ Tcji[:3, 3] + pj).unsqueeze(-1)).squeeze(-1) + Tpj[:3, 3]
        if b.parent >= 0:
            Rp, pp = Rs[b.parent], ps[b.parent]
            p = (Rp @ p.unsqueeze(-1)).squeeze(-1) + pp
            R = Rp @ R
        Rs.append(R)
        ps.append(p)
    return torch.stack(Rs, 1), torch.stack(ps, 1)


def body_point_world(model: Model, q: torch.Tensor, body: int, local=(0.0, 0.0, 0.0)) -> torch.Tensor:
    """bodynodes[body].to_world(local) -> [N,3]."""
    R, p = body_transforms(model, q)
    loc = torch.tensor(np.asarray(local, dtype=np.float64), device=q.device)
    return (R[:, body] @ loc).reshape(q.shape[0], 3) + p[:, body]


def body_com_world(model: Model, q: torch.Tensor) -> torch.Tensor:
    """bodynodes[*].com() -> [N,nb,3]."""
    R, p = body_transforms(model, q)
    loc = torch.tensor(np.array([b.com for b in model.bodies], dtype=np.float64), device=q.device)   # [nb,3]
    return (R @ loc[None, :, :, None]).squeeze(-1) + p


def body_com_spatial_velocities(model: Model, q: torch.Tensor, dq: torch.Tensor) -> torch.Tensor:
    """bodynodes[*].com_spatial_velocity() -> [N,nb,6] = [angular; linear velocity of the COM] relative to the world,
    expressed in BODY coordinates (pydart2's no-argument call is DART's BodyNode::getCOMSpatialVelocity())."""
    n, dev, dt = q.shape[0], q.device, torch.float64
    q, dq = q.to(dt), dq.to(dt)
    R, p = body_transforms(model, q)
    com = body_com_world(model, q)
    W, V = [], []          # angular velocity and linear velocity of the body ORIGIN
    for k, b in enumerate(model.bodies):
        if b.parent >= 0:
            wp, vp = W[b.parent], V[b.parent] + torch.cross(W[b.parent], p[:, k] - p[:, b.parent], dim=1)
            Rp = R[:, b.parent]
        else:
            wp = torch.zeros((n, 3), dtype=dt, device=dev); vp = torch.zeros((n, 3), dtype=dt, device=dev)
            Rp = torch.eye(3, dtype=dt, device=dev).expand(n, 3, 3)
        w, v = wp, vp
        if b.dof >= 0:
            Tpj = torch.tensor(b.T_parent_joint, dtype=dt, device=dev)
            axis_w = (Rp @ (Tpj[:3, :3] @ torch.tensor(b.axis, dtype=dt, device=dev)))          # joint axis, world axes
            if b.joint_type == JOINT_REVOLUTE:
                # the joint frame origin in the world: parent origin + Rp * Tpj.p ; the body origin turns about it
                jo = p[:, b.parent] + (Rp @ Tpj[:3, 3]) if b.parent >= 0 else Tpj[:3, 3].expand(n, 3)
                wj = axis_w * dq[:, b.dof, None]
                w = wp + wj
                v = vp + torch.cross(wj, p[:, k] - jo, dim=1)
            elif b.joint_type == JOINT_PRISMATIC:
                v = vp + axis_w * dq[:, b.dof, None]
        W.append(w); V.append(v)
    W, V = torch.stack(W, 1), torch.stack(V, 1)
    vc = V + torch.cross(W, com - p, dim=2)
    Rt = R.transpose(-1, -2)
    return torch.cat([(Rt @ W[..., None]).squeeze(-1), (Rt @ vc[..., None]).squeeze(-1)], dim=2)
