"""TEST INFRASTRUCTURE: numpy interpreter of planner PlanOps (qibojit_b200/planner.py), used to
check the pass partition / round schedule / diagonal merging on the CPU against gate-by-gate
application."""

import numpy as np


def apply_planop(state, op, nqubits):
    idx = np.arange(state.size)
    cmask = 0
    for c in op.controls:
        cmask |= 1 << c
    active = (idx & cmask) == cmask
    if op.kind == "diag":
        t = np.zeros_like(idx)
        for j, b in enumerate(op.targets):
            t |= ((idx >> b) & 1) << j
        table = np.asarray(op.data, dtype=np.complex128).reshape(-1)
        out = state.copy()
        out[active] = state[active] * table[t[active]]
        return out
    assert op.kind == "dense"
    k = len(op.targets)
    u = np.asarray(op.data, dtype=np.complex128).reshape(1 << k, 1 << k)
    tmask = 0
    for b in op.targets:
        tmask |= 1 << b
    base = idx[active & ((idx & tmask) == 0)]
    offs = [sum(((e >> j) & 1) << op.targets[j] for j in range(k)) for e in range(1 << k)]
    x = np.stack([state[base + o] for o in offs])   # (2^k, ngroups)
    y = u @ x
    out = state.copy()
    for e, o in enumerate(offs):
        out[base + o] = y[e]
    return out


def run_plan(state, plan, nqubits, apply_raw, nreg=None, fixed=()):
    """plan: output of planner.plan_queue; apply_raw(state, gate) handles raw gates."""
    for seg in plan:
        if seg[0] == "raw":
            state = apply_raw(state, seg[1])
            continue
        local = set(seg[1])
        for regs, ops in seg[2]:
            assert set(regs) <= local and list(regs) == sorted(set(regs)), "register bits must be local bits"
            assert set(fixed) <= set(regs)
            if nreg is not None:
                assert len(regs) == nreg
            for op in ops:
                if op.kind == "dense":
                    assert set(op.targets) <= set(regs), "dense target outside the round's register bits"
                state = apply_planop(state, op, nqubits)
    return state
