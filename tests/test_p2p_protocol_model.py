"""Executable model of the peer-memory protocol of wl_p2p.cuh / wl_dist.cu / wl_vcycle.cu, on CPU.

The CUDA path itself needs ≥2 GPUs (tests/test_multi_gpu.py, tests/dist_check.py).  What can be
checked without them is the LOGIC of the epoch protocol: R threads play the ranks, plain Python
lists play the flag block and ghost planes of every rank's arena, and random delays shuffle the
interleavings.  The model follows the kernels statement by statement:

 * halo push (k_p2p_push / push_segments): post ready(e) → wait ready from the peers → store the
   planes into the peers' ghosts → post done(e) → wait done from the peers → consume; epochs are
   monotonic and compared with >=; the persistent solver kernel posts ready(e+1) EARLY, right after
   the consumer phase (post_ready), the stand-alone push kernel posts it at its own start;
 * all-reduce (p2p_allreduce2 / reduce<>): every rank stores {value, epoch} into slot [epoch & 1][me]
   of every rank's mailbox, waits until its own slots hold the epoch, adds in rank order.

Invariants asserted: a ghost plane is never overwritten while its owner may still read it, every
consumer sees exactly the neighbour's data of the current epoch, two mailbox slots suffice, and all
ranks obtain bit-identical sums."""
import random
import threading
import time

import pytest


class Arena:
    def __init__(self, R):
        self.ready = [0] * R
        self.done = [0] * R
        self.mbox = [[(None, 0)] * R for _ in range(2)]
        self.ghost = {"lo": None, "hi": None}
        self.reading = False      # the owner is inside a phase that reads its ghost planes


def _spin(cond, deadline):
    while not cond():
        if time.monotonic() > deadline:
            raise TimeoutError("protocol deadlock")
        time.sleep(0)


def _rank(me, R, arenas, steps, early_post, errors, seed):
    rng = random.Random(seed)
    mine = arenas[me]
    up, dn = (me + 1) % R, (me - 1) % R
    peers = sorted({up, dn} - {me})
    push_epoch = ar_epoch = 0
    deadline = time.monotonic() + 20
    try:
        if early_post:                      # kernel start: every earlier kernel of this rank has retired
            for p in peers:
                arenas[p].ready[me] = push_epoch + 1
        for step in range(steps):
            # ---- producer phase, then the push -------------------------------------------------
            payload = (me, step)
            push_epoch += 1
            e = push_epoch
            if not early_post:
                for p in peers:
                    arenas[p].ready[me] = e
            _spin(lambda: all(mine.ready[p] >= e for p in peers), deadline)
            if rng.random() < 0.3:
                time.sleep(rng.random() * 1e-4)
            for p, side in ((up, "lo"), (dn, "hi")):   # my top planes → lower ghosts of `up`, bottom → upper ghosts of `dn`
                assert not arenas[p].reading, f"rank {me} overwrote ghosts rank {p} was still reading (epoch {e})"
                arenas[p].ghost[side] = payload
            for p in peers:
                arenas[p].done[me] = e
            _spin(lambda: all(mine.done[p] >= e for p in peers), deadline)
            # ---- consumer phase -------------------------------------------------------------------
            mine.reading = True
            if rng.random() < 0.3:
                time.sleep(rng.random() * 1e-4)
            assert mine.ghost["lo"] == (dn, step), (me, step, mine.ghost)
            assert mine.ghost["hi"] == (up, step), (me, step, mine.ghost)
            mine.reading = False
            if early_post:                  # post_ready: reads of pushes ≤ e are complete
                for p in peers:
                    arenas[p].ready[me] = push_epoch + 1
            # ---- all-reduce of a value that depends on the rank and the step -----------------------
            ar_epoch += 1
            a = ar_epoch
            slot = a & 1
            val = float((me + 1) * 1e-3 + step)
            for r in range(R):
                arenas[r].mbox[slot][me] = (val, a)
            tot = 0.0
            for r in range(R):
                _spin(lambda r=r: mine.mbox[slot][r][1] == a, deadline)
                tot += mine.mbox[slot][r][0]
            expect = 0.0
            for r in range(R):
                expect += float((r + 1) * 1e-3 + step)
            assert tot == expect, (me, step, tot, expect)   # same order on every rank → same bits
            if rng.random() < 0.2:
                time.sleep(rng.random() * 2e-4)             # let the others run ahead
    except BaseException as ex:  # noqa: BLE001 - reported to the main thread
        errors.append((me, repr(ex)))


@pytest.mark.parametrize("R", [2, 3, 8])
@pytest.mark.parametrize("early_post", [False, True])
def test_epoch_protocol_model(R, early_post):
    arenas = [Arena(R) for _ in range(R)]
    errors = []
    threads = [threading.Thread(target=_rank, args=(r, R, arenas, 150, early_post, errors, 1000 * R + r)) for r in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not errors, errors
    assert all(not t.is_alive() for t in threads)


def test_one_mailbox_slot_is_not_enough():
    """The reason for the parity double-buffering: with ONE slot a fast rank's epoch e+1 mail can
    overwrite its epoch-e mail before a slow rank has read it."""
    mbox = [(None, 0)] * 2            # rank 1's mailbox, single slot per source
    mbox[0] = (10.0, 1)               # rank 0 mails epoch 1
    # rank 0 finishes epoch 1 as soon as rank 1's mail arrives and immediately mails epoch 2 …
    mbox[0] = (20.0, 2)
    # … while rank 1 is still polling for epoch 1: it never sees it
    assert mbox[0][1] != 1
    # with two slots, epoch 2 goes to the other slot and epoch 3 needs rank 1's contribution to epoch 2 first
    slots = [[(None, 0)] * 2 for _ in range(2)]
    slots[1][0] = (10.0, 1)
    slots[0][0] = (20.0, 2)
    assert slots[1][0] == (10.0, 1)
