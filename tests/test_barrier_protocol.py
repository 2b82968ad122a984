"""Exhaustive interleaving check of the cross-GPU flag protocol used by K-AR (csrc/comm_ar.cu,
``block_barrier_all_gpus``) - the "race detector" for the part compute-sanitizer cannot see.

Model of one block index on W ranks, E consecutive kernel launches (epochs):

    write own gradients (version e)                      <- GEMM epilogues of step e
    barrier A:  flag[peer][A][me] = e  for every peer;   wait flag[me][A][peer] >= e
    read every peer's gradients                          <- must observe exactly version e
    barrier B:  flag[peer][B][me] = e;                   wait flag[me][B][peer] >= e
    (next launch overwrites / zeroes the gradients)      <- must not happen while a peer still reads version e

Flags are never reset; epochs only grow, so a stale flag can never satisfy a wait.  Every store and every
poll is one atomic step; the checker explores ALL interleavings (DFS with memoisation) and asserts there is no
deadlock and no read of a wrong gradient version.  A deliberately broken variant (one shared slot for both phases)
must be caught - that keeps the checker honest.
"""
import itertools

import pytest


def _program(world: int, me: int, epochs: int, shared_phase_slot: bool):
    ops = []
    peers = [p for p in range(world) if p != me]
    for e in range(1, epochs + 1):
        ops.append(("write_grad", e))
        for ph in (0, 1):
            slot_ph = 0 if shared_phase_slot else ph
            for p in peers:
                ops.append(("set_flag", p, slot_ph, e))
            for p in peers:
                ops.append(("wait_flag", p, slot_ph, e))
            if ph == 0:
                for p in peers:
                    ops.append(("read_grad", p, e))
    return ops


def _explore(world: int, epochs: int, shared_phase_slot: bool = False):
    progs = [_program(world, r, epochs, shared_phase_slot) for r in range(world)]
    # state: (pc per rank, flags[(dst, phase, src)] -> epoch, grad version per rank)
    flag_keys = [(d, ph, s) for d in range(world) for ph in (0, 1) for s in range(world) if s != d]
    start = (tuple(0 for _ in range(world)), tuple(0 for _ in flag_keys), tuple(0 for _ in range(world)))
    kidx = {k: i for i, k in enumerate(flag_keys)}
    seen, stack = {start}, [start]
    violations, deadlocks, finals = [], 0, 0
    while stack:
        pcs, flags, grads = stack.pop()
        if all(pcs[r] == len(progs[r]) for r in range(world)):
            finals += 1
            continue
        moved = False
        for r in range(world):
            if pcs[r] == len(progs[r]):
                continue
            op = progs[r][pcs[r]]
            nflags, ngrads = flags, grads
            if op[0] == "write_grad":
                ngrads = grads[:r] + (op[1],) + grads[r + 1:]
            elif op[0] == "set_flag":
                i = kidx[(op[1], op[2], r)]
                nflags = flags[:i] + (max(flags[i], op[3]),) + flags[i + 1:]
            elif op[0] == "wait_flag":
                if flags[kidx[(r, op[2], op[1])]] < op[3]:
                    continue                                      # still spinning: not enabled
            elif op[0] == "read_grad":
                if grads[op[1]] != op[2]:
                    violations.append((r, op, grads))
            moved = True
            npcs = pcs[:r] + (pcs[r] + 1,) + pcs[r + 1:]
            nxt = (npcs, nflags, ngrads)
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not moved:
            deadlocks += 1
    return dict(states=len(seen), violations=violations, deadlocks=deadlocks, finals=finals)


@pytest.mark.parametrize("world,epochs", [(2, 3), (3, 2)])
def test_two_phase_epoch_barrier_is_safe_and_live(world, epochs):
    r = _explore(world, epochs)
    assert r["deadlocks"] == 0 and r["finals"] >= 1
    assert not r["violations"], r["violations"][:3]
    assert r["states"] > 50                                        # the exploration really branched


def test_checker_catches_a_broken_protocol():
    # one slot for both phases: a fast rank's phase-B flag of epoch e satisfies a slow rank's ... nothing yet, but its
    # phase-A flag of epoch e+1 is indistinguishable from phase B of e+1 -> a peer can leave barrier B(e+1) early and
    # overwrite gradients that are still being read
    r = _explore(2, 3, shared_phase_slot=True)
    assert r["violations"] or r["deadlocks"]


# ------------------------------------------------------------------------------------------------ one-shot push exchange
def _explore_push(world: int, epochs: int, n_slots: int = 2):
    """The push / LL-style exchange (csrc/comm_ar.cu allreduce_push_update_kernel, SPARKTORCH_EXCHANGE=push): the GEMM epilogues of
    step e mirror the rank's gradients into its receive slot (parity e % n_slots) on EVERY peer, ONE flag exchange says "my pushes
    of step e are complete", then every operand of the reduction is local.  There is no second barrier: what protects a slot that
    a slow rank is still reading from the pushes of step e + n_slots is that a rank reaches that step only after the flag exchange
    of step e + 1, which the slow rank enters after its reads.  The checker explores all interleavings; with a single slot
    (n_slots = 1) the same program must be caught overwriting live data."""
    progs = []
    for me in range(world):
        peers = [p for p in range(world) if p != me]
        ops = []
        for e in range(1, epochs + 1):
            for p in peers:
                ops.append(("push", p, e % n_slots, e))          # slot[p][parity][me] = e
            for p in peers:
                ops.append(("set_flag", p, e))
            for p in peers:
                ops.append(("wait_flag", p, e))
            for p in peers:
                ops.append(("read_slot", p, e % n_slots, e))      # slot[me][parity][p] must hold e
        progs.append(ops)
    pairs = [(d, s) for d in range(world) for s in range(world) if s != d]
    fidx = {k: i for i, k in enumerate(pairs)}
    skeys = [(d, par, s) for d in range(world) for par in range(n_slots) for s in range(world) if s != d]
    sidx = {k: i for i, k in enumerate(skeys)}
    start = (tuple(0 for _ in range(world)), tuple(0 for _ in pairs), tuple(0 for _ in skeys))
    seen, stack = {start}, [start]
    violations, deadlocks, finals = [], 0, 0
    while stack:
        pcs, flags, slots = stack.pop()
        if all(pcs[r] == len(progs[r]) for r in range(world)):
            finals += 1
            continue
        moved = False
        for r in range(world):
            if pcs[r] == len(progs[r]):
                continue
            op = progs[r][pcs[r]]
            nflags, nslots = flags, slots
            if op[0] == "push":
                i = sidx[(op[1], op[2], r)]
                nslots = slots[:i] + (op[3],) + slots[i + 1:]
            elif op[0] == "set_flag":
                i = fidx[(op[1], r)]
                nflags = flags[:i] + (max(flags[i], op[2]),) + flags[i + 1:]
            elif op[0] == "wait_flag":
                if flags[fidx[(r, op[1])]] < op[2]:
                    continue
            elif op[0] == "read_slot":
                if slots[sidx[(r, op[2], op[1])]] != op[3]:
                    violations.append((r, op))
            moved = True
            nxt = (pcs[:r] + (pcs[r] + 1,) + pcs[r + 1:], nflags, nslots)
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not moved:
            deadlocks += 1
    return dict(states=len(seen), violations=violations, deadlocks=deadlocks, finals=finals)


@pytest.mark.parametrize("world,epochs", [(2, 4), (3, 3)])
def test_one_shot_push_exchange_needs_no_second_barrier(world, epochs):
    r = _explore_push(world, epochs, n_slots=2)
    assert r["deadlocks"] == 0 and r["finals"] >= 1 and r["states"] > 50
    assert not r["violations"], r["violations"][:3]


def test_push_exchange_with_a_single_receive_slot_is_caught():
    r = _explore_push(2, 3, n_slots=1)
    assert r["violations"], "a fast rank's pushes of step e+1 must be seen overwriting the slot a slow rank still reads"
