"""Model check of the NVLink mailbox protocol of csrc/comm.cu / csr_tile.cu under random
schedules: mailboxes double-buffered by the parity of the exchange number, per-source sequence flags, pushes that never
wait and no acknowledgement travelling back.

Every rank runs the program of the fused compute + exchange kernel, exchange after exchange:
    PUSH(k):  store this rank's boundary values into box[k & 1][me] of every neighbour, then raise flag[k & 1][me] = k there
    READ(k):  wait until flag[k & 1][nb] >= k for every neighbour nb, then read box[k & 1][nb]
PUSH(k + 1) follows READ(k) in program order (the next kernel starts after this one has ended).  The scheduler picks
any rank whose next step can run.  Checked for chains and rings of ranks, with and without a periodic all-reduce:
    * no deadlock: some rank can always run until all have finished;
    * every READ(k) sees the values of exchange k (a box is never overwritten before it has been consumed), although
      nobody ever tells a sender that its previous message has been read;
    * a rank is never more than one exchange ahead of a neighbour.
Pure Python, no GPU: the hardware counterpart is the burst test of tests/dist_gpu_worker.py."""
import random

import pytest


def run(n_ranks, ring, n_exchanges, seed, allreduce_every=0):
    rng = random.Random(seed)
    nbrs = []
    for r in range(n_ranks):
        nb = [q for q in (r - 1, r + 1) if 0 <= q < n_ranks]
        if ring and n_ranks > 2:
            nb = [(r - 1) % n_ranks, (r + 1) % n_ranks]
        nbrs.append(sorted(set(nb)))
    box = [[[None] * n_ranks for _ in range(2)] for _ in range(n_ranks)]   # box[dst][parity][src] = (k, payload)
    flag = [[[0] * n_ranks for _ in range(2)] for _ in range(n_ranks)]
    ar_arrived = [0] * n_ranks                                             # all-reduces each rank has entered
    # program counter: (k, phase) with phase 0 = PUSH(k), 1 = READ(k), 2 = all-reduce after exchange k (optional)
    pc = [(1, 0)] * n_ranks
    done = [False] * n_ranks
    max_lead = 0
    steps = 0
    while not all(done):
        runnable = []
        for r in range(n_ranks):
            if done[r]:
                continue
            k, ph = pc[r]
            if ph == 0:
                runnable.append(r)                                          # pushes never wait
            elif ph == 1:
                if all(flag[r][k & 1][q] >= k for q in nbrs[r]):
                    runnable.append(r)
            else:
                # enter the all-reduce at once; leave it when everybody has entered this one
                runnable.append(r)
        assert runnable, f"deadlock at {pc}"
        r = rng.choice(runnable)
        k, ph = pc[r]
        if ph == 0:
            for q in nbrs[r]:
                box[q][k & 1][r] = (k, (r, k))
                flag[q][k & 1][r] = k
            pc[r] = (k, 1)
        elif ph == 1:
            for q in nbrs[r]:
                got = box[r][k & 1][q]
                assert got == (k, (q, k)), f"rank {r} exchange {k}: read {got} from {q}"
                k_q = pc[q][0] if not done[q] else n_exchanges
                max_lead = max(max_lead, abs(k_q - k))
            if allreduce_every and k % allreduce_every == 0:
                ar_arrived[r] += 1
                pc[r] = (k, 2)
            elif k == n_exchanges:
                done[r] = True
            else:
                pc[r] = (k + 1, 0)
        else:
            if min(ar_arrived) >= ar_arrived[r]:                            # everybody has entered this all-reduce
                if k == n_exchanges:
                    done[r] = True
                else:
                    pc[r] = (k + 1, 0)
        steps += 1
        assert steps < 200 * n_ranks * n_exchanges, "livelock"
    return max_lead


@pytest.mark.parametrize("n_ranks,ring", [(2, False), (3, False), (8, False), (4, True), (5, True)])
@pytest.mark.parametrize("allreduce_every", [0, 2])
def test_mailbox_protocol_random_schedules(n_ranks, ring, allreduce_every):
    worst = 0
    for seed in range(60):
        worst = max(worst, run(n_ranks, ring, 24, seed, allreduce_every))
    assert worst <= 1     # a rank is never more than one exchange ahead of a neighbour


def test_model_detects_a_single_buffer():
    """The same program with ONE mailbox per source (no parity) loses messages under some schedule: the model is able
    to see the failure the double buffer prevents."""
    def run_single(seed):
        rng = random.Random(seed)
        box = [[None, None], [None, None]]       # box[dst][src]
        flag = [[0, 0], [0, 0]]
        pc = [(1, 0), (1, 0)]
        done = [False, False]
        while not all(done):
            runnable = [r for r in range(2) if not done[r] and (pc[r][1] == 0 or flag[r][1 - r] >= pc[r][0])]
            r = rng.choice(runnable)
            k, ph = pc[r]
            q = 1 - r
            if ph == 0:
                box[q][r] = k
                flag[q][r] = k
                pc[r] = (k, 1)
            else:
                if box[r][q] != k:
                    return False
                if k == 12:
                    done[r] = True
                else:
                    pc[r] = (k + 1, 0)
        return True

    assert not all(run_single(s) for s in range(200))
