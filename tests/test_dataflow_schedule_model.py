"""CPU: the arithmetic the dataflow replay scheduler (lightfm_b200/csrc/lfm_replay_dataflow.cuh) relies
on, restated in numpy / plain Python and checked against the sequential definitions it replaces:

  * rand_r's state after k draws is A^k * s + C_k (the per-lane jump-ahead table);
  * draw % n by a multiply-high with ceil(2^64 / n) (Lemire 2019) equals the division;
  * judging W draws per interaction and resolving the lanes in order (incoming shift -> own
    rejections -> outgoing shift, with the reference's give-up rule T:1123-1127) consumes the draw
    stream exactly as the one-draw-at-a-time loop of fit_bpr does;
  * the chunk-wise version count (counters + earlier lanes of the chunk, one store per row by its
    last lane) equals counting earlier touches one interaction at a time.

The CUDA code itself is checked bit for bit on the GPU (tests/test_gpu_replay_dataflow.py)."""
import numpy as np

A, C, MASK = 1103515245, 12345, 0xFFFFFFFF


def _temper(x):
    x ^= x >> 11
    x ^= (x << 7) & 0x9D2C5680
    x ^= (x << 15) & 0xEFC60000
    x ^= x >> 18
    return x & MASK


def _rand_r(seed):  # T:64-81
    seed = (seed * A + C) & MASK
    return seed, _temper(seed) >> 1


def test_lcg_jump_ahead_table():
    ja, jc = [A], [C]
    for _ in range(32):
        jc.append((jc[-1] * A + C) & MASK)
        ja.append((ja[-1] * A) & MASK)
    rng = np.random.default_rng(0)
    for s0 in rng.integers(0, 2 ** 32, 50):
        s = int(s0)
        for k in range(1, 33):
            s, _ = _rand_r(s)
            assert s == (ja[k - 1] * int(s0) + jc[k - 1]) & MASK


def test_multiply_high_modulo():
    rng = np.random.default_rng(1)
    for n in [1, 2, 3, 7, 100_000, 19_800_000, 2 ** 31 - 1] + [int(x) for x in rng.integers(1, 2 ** 31, 200)]:
        magic = (2 ** 64 - 1) // n + 1
        for r in [0, 1, n - 1, n, n + 1, 2 ** 31 - 1] + [int(x) for x in rng.integers(0, 2 ** 31, 200)]:
            low = (magic * r) & (2 ** 64 - 1)
            assert (low * n) >> 64 == r % n


def _sequential(users, member, n_total, draws):
    """fit_bpr's loop: per interaction draw until the candidate is not a positive of the user, give
    up after n_total draws and keep the last one.  Returns (accepted draw index per interaction,
    draws consumed)."""
    q, out = 0, []
    for u in users:
        for j in range(n_total):
            d = q
            q += 1
            if not member[u][draws[d]]:
                break
        out.append(d)
    return out, q


def _windowed(users, member, n_total, draws, W):
    """The scheduler's rounds: every unresolved interaction judges W consecutive draws, then the
    lanes are resolved in order."""
    out = [None] * len(users)
    qbase, start, attempts = 0, 0, 0
    while start < len(users):
        sft, over = 0, None
        for r in range(start, len(users)):
            k0 = r - start
            lim = max(0, min(64, n_total - 1 - (attempts if r == start else 0)))
            run = 0
            while sft + run < W and member[users[r]][draws[qbase + k0 + sft + run]]:
                run += 1
            rr = min(run, lim)
            if sft + rr > W - 1:
                over = (r, sft)
                break
            out[r] = qbase + k0 + sft + rr
            sft += rr
        if over is None:
            qbase += (len(users) - start) + sft
            start = len(users)
        else:
            r, s_in = over
            qbase += (r - start) + W
            attempts = (attempts if r == start else 0) + (W - s_in)
            start = r
    return out, qbase


def test_windowed_resolution_equals_one_draw_at_a_time():
    rng = np.random.default_rng(2)
    for trial in range(300):
        n_users, n_items = int(rng.integers(1, 6)), int(rng.integers(1, 9))
        density = rng.choice([0.0, 0.2, 0.6, 0.95, 1.0])
        member = rng.random((n_users, n_items)) < density
        n_inter = int(rng.integers(1, 33))
        n_total = int(rng.choice([1, 2, 3, n_inter, 50]))       # no_examples: the give-up bound
        users = rng.integers(0, n_users, n_inter)
        draws = rng.integers(0, n_items, 64 * 64)
        want, q_want = _sequential(users, member, n_total, draws)
        for W in (1, 2, 8):
            got, q_got = _windowed(users, member, n_total, draws, W)
            assert got == want and q_got == q_want, (trial, W)


def test_chunked_version_count_equals_sequential():
    rng = np.random.default_rng(3)
    for trial in range(100):
        n_users, n_items, n = int(rng.integers(1, 40)), int(rng.integers(2, 40)), int(rng.integers(1, 200))
        user = rng.integers(0, n_users, n)
        item = rng.integers(0, n_items, n)
        neg = rng.integers(0, n_items, n)
        neg[neg == item] = -1                                     # the scheduler's "negative = positive" marker
        # sequential definition: version = touches of the row by earlier interactions
        cu, ci = np.zeros(n_users, int), np.zeros(n_items, int)
        want = []
        for t in range(n):
            want.append((cu[user[t]], ci[item[t]], ci[neg[t]] if neg[t] >= 0 else 0))
            cu[user[t]] += 1
            ci[item[t]] += 1
            if neg[t] >= 0:
                ci[neg[t]] += 1
        # chunk-wise: counters + earlier lanes of the chunk; the last lane on a row stores the new count
        cu[:], ci[:] = 0, 0
        got = []
        for t0 in range(0, n, 32):
            lanes = range(t0, min(n, t0 + 32))
            vers, stores_u, stores_i = [], {}, {}
            for l in lanes:
                eu = cu[user[l]] + sum(user[j] == user[l] for j in lanes if j < l)
                ei = ci[item[l]] + sum((item[j] == item[l]) + (neg[j] == item[l]) for j in lanes if j < l)
                en = 0
                if neg[l] >= 0:
                    en = ci[neg[l]] + sum((item[j] == neg[l]) + (neg[j] == neg[l]) for j in lanes if j < l)
                vers.append((eu, ei, en))
                stores_u[user[l]] = eu + 1                        # later lanes overwrite: the last lane wins
                stores_i[item[l]] = ei + 1
                if neg[l] >= 0:
                    stores_i[neg[l]] = en + 1
            for k, v in stores_u.items():
                cu[k] = v
            for k, v in stores_i.items():
                ci[k] = v
            got += vers
        assert got == want, trial
