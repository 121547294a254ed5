"""Second, literal-loop restatement of the quirk-bearing pieces of the hot path.
TEST INFRASTRUCTURE ONLY (see oracle/deepj_oracle.py header; parity unpinned).

Plain NumPy / Python loops written independently of the vectorised torch
oracle, following the reference line by line, so that the two restatements pin
each other (tests/test_oracle.py)."""
import numpy as np


def pitch_bins_tf_emulation(x, octave=12, num_octaves=4):
    """model.py:43-49 op by op with NumPy equivalents of the TF calls:
    reduce_sum(list, axis=3) -> tile([4,1,1]) -> reshape([B,T,48,1])."""
    B, T = x.shape[0], x.shape[1]
    stacked = np.stack([x[:, :, i::octave, 0] for i in range(octave)])     # tf packs the list: [12,B,T,4]
    bins = stacked.sum(axis=3)                                              # [12,B,T]
    bins = np.tile(bins, [num_octaves, 1, 1])                               # [48,B,T]
    return np.reshape(bins, [B, T, octave * num_octaves, 1])


def pitch_bins_closed_form(x, octave=12, num_octaves=4):
    """The index formula of SURVEY.md 8a/A7, as loops."""
    B, T, N = x.shape[0], x.shape[1], octave * num_octaves
    out = np.zeros((B, T, N, 1), dtype=x.dtype)
    BT = B * T
    for b in range(B):
        for t in range(T):
            for n in range(N):
                flat = (b * T + t) * N + n
                p, r = flat // BT, flat % BT
                i, b2, t2 = p % octave, r // T, r % T
                out[b, t, n, 0] = sum(x[b2, t2, octave * o + i, 0] for o in range(num_octaves))
    return out


def conv1d_same_loops(x, W, b):
    """Conv1D(padding='same') over notes, TF SAME for even k: 11 left / 12 right
    zeros, cross-correlation (model.py:56)."""
    R, N, Cc = x.shape
    k, _, Oc = W.shape
    left = (k - 1) // 2
    out = np.zeros((R, N, Oc))
    for r in range(R):
        for n in range(N):
            acc = b.astype(np.float64).copy()
            for kk in range(k):
                src = n - left + kk
                if 0 <= src < N:
                    acc += x[r, src] @ W[kk]
            out[r, n] = acc
    return out


def lstm_loops(x, W, U, b, hard=True):
    """keras LSTM, gate order i,f,c,o, zero initial state."""
    S, steps, _ = x.shape
    u = U.shape[0]
    act = (lambda z: np.clip(0.2 * z + 0.5, 0, 1)) if hard else (lambda z: 1 / (1 + np.exp(-z)))
    out = np.zeros((S, steps, u))
    for s in range(S):
        h = np.zeros(u); c = np.zeros(u)
        for t in range(steps):
            z = x[s, t] @ W + h @ U + b
            i, f, g, o = act(z[:u]), act(z[u:2 * u]), np.tanh(z[2 * u:3 * u]), act(z[3 * u:])
            c = f * c + i * g
            h = o * np.tanh(c)
            out[s, t] = h
    return out


def primary_loss_loops(y_true, y_pred, eps=1e-7):
    """model.py:14-20 element by element (Keras clip -> logit -> sigmoid CE)."""
    B, T, N, _ = y_true.shape

    def bce(t, o):
        o = min(max(o, eps), 1 - eps)
        x = np.log(o / (1 - o))
        return max(x, 0) - x * t + np.log1p(np.exp(-abs(x)))
    tot = 0.0
    for b in range(B):
        for t in range(T):
            a = bb = c = 0.0
            for n in range(N):
                yt, yp = y_true[b, t, n], y_pred[b, t, n]
                played = yt[0]
                a += bce(yt[0], yp[0])
                bb += bce(yt[1], played * yp[1] + (1 - played) * yt[1])
                c += (yt[2] - (played * yp[2] + (1 - played) * yt[2])) ** 2
            tot += (a + bb + c) / N
    return tot / (B * T)


def nadam_scalar_loop(p, grads_seq, lr=0.002, b1=0.9, b2=0.999, eps=1e-8, sd=0.004):
    """keras.optimizers.Nadam.get_updates for one scalar parameter.  lr / beta_1 / beta_2 are float32 K.variables in
    Keras: the update is a function of their float32-rounded values."""
    lr, b1, b2 = (float(np.float32(x)) for x in (lr, b1, b2))
    m = v = 0.0
    m_schedule = 1.0
    for it, g in enumerate(grads_seq):
        t = it + 1
        mu_t = b1 * (1 - 0.5 * 0.96 ** (t * sd))
        mu_t1 = b1 * (1 - 0.5 * 0.96 ** ((t + 1) * sd))
        ms_new = m_schedule * mu_t
        ms_next = m_schedule * mu_t * mu_t1
        g_prime = g / (1 - ms_new)
        m = b1 * m + (1 - b1) * g
        m_prime = m / (1 - ms_next)
        v = b2 * v + (1 - b2) * g * g
        v_prime = v / (1 - b2 ** t)
        m_bar = (1 - mu_t) * g_prime + mu_t1 * m_prime
        p = p - lr * m_bar / (np.sqrt(v_prime) + eps)
        m_schedule = ms_new
    return p
