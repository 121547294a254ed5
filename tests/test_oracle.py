"""CPU tests: pin the oracle (the reference ships no golden vectors for
model.py / generate.py -- SURVEY.md 4, 8c) with hand-computed cases for every
quirk of SURVEY.md 9/H1 and with an independent literal-loop restatement."""
import os

import numpy as np
import torch

from oracle import deepj_oracle as O
from oracle import literal as L
import helpers

CFG = O.Config()
GOLD = os.path.join(os.path.dirname(__file__), "golden", "deepj_small.npz")


def test_param_inventory():
    shp = O.param_shapes(CFG)
    assert len(shp) == 28
    assert sum(int(np.prod(s)) for s in shp.values()) == 1269476
    p = O.init_params(CFG, 0)
    u = CFG.time_axis_units
    assert torch.all(p["time0.lstm.b"][u:2 * u] == 1) and p["time0.lstm.b"].sum() == u   # unit_forget_bias


def test_q1_pitch_bins_scramble():
    rs = np.random.RandomState(0)
    for B, T in ((1, 4), (3, 5), (2, 128)):
        x = rs.rand(B, T, 48, 3)
        ref = L.pitch_bins_tf_emulation(x)
        if B * T <= 32:
            np.testing.assert_allclose(L.pitch_bins_closed_form(x), ref, rtol=0, atol=1e-12)
        got = O.pitch_bins(torch.tensor(x), CFG).numpy()
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    # the feature really mixes batch elements: changing sequence 1 changes sequence 0's bins
    x = rs.rand(2, 4, 48, 3); y = x.copy(); y[1] += 1
    a, b = O.pitch_bins(torch.tensor(x), CFG), O.pitch_bins(torch.tensor(y), CFG)
    assert not torch.allclose(a[0], b[0])


def test_q3_conv_same_padding_11_12():
    rs = np.random.RandomState(1)
    x, W, b = rs.randn(2, 48, 3), rs.randn(24, 3, 5), rs.randn(5)
    got = O.conv1d_same(torch.tensor(x), torch.tensor(W), torch.tensor(b)).numpy()
    np.testing.assert_allclose(got, L.conv1d_same_loops(x, W, b), atol=1e-10)
    # impulse at note 0 with a one-hot kernel tap k reaches output note 11-k (left pad 11)
    x = np.zeros((1, 48, 3)); x[0, 0, 0] = 1
    W = np.zeros((24, 3, 1)); W[3, 0, 0] = 1
    out = O.conv1d_same(torch.tensor(x), torch.tensor(W), torch.zeros(1, dtype=torch.float64)).numpy()[0, :, 0]
    assert out[11 - 3] == 1 and out.sum() == 1


def test_q2_lstm_hard_sigmoid_and_gate_order():
    rs = np.random.RandomState(2)
    x, W, U, b = rs.randn(3, 6, 5), rs.randn(5, 16), rs.randn(4, 16), rs.randn(16)
    t = lambda a: torch.tensor(a)
    for hard, name in ((True, "hard_sigmoid"), (False, "sigmoid")):
        got = O.lstm_seq(t(x), t(W), t(U), t(b), name).numpy()
        np.testing.assert_allclose(got, L.lstm_loops(x, W, U, b, hard), atol=1e-10)
    assert O.hard_sigmoid(torch.tensor([-3.0, -2.5, 0.0, 1.0, 2.5, 9.0], dtype=torch.float64)).tolist() == [0, 0, 0.5, 0.7, 1, 1]


def test_q4_shift_chosen_three_channels():
    p = O.init_params(CFG, 0, torch.float64)
    rs = np.random.RandomState(3)
    B, T = 1, 2
    time_out = torch.zeros(B, T, 48, CFG.time_axis_units, dtype=torch.float64)
    chosen = torch.tensor(rs.rand(B, T, 48, 3))
    style = torch.zeros(B, T, CFG.style_units, dtype=torch.float64)
    taps = {}
    O.note_axis_forward(p, CFG, time_out, chosen, style, taps=taps)
    x = taps["note0.in"]
    assert x.shape[-1] == CFG.time_axis_units + 3
    sp = torch.tanh(p["note0.sd.b"])
    np.testing.assert_allclose((x[0, 0, 5, -3:] - sp[-3:]).numpy(), chosen[0, 0, 4].numpy(), atol=1e-12)
    np.testing.assert_allclose((x[0, 0, 0, -3:] - sp[-3:]).numpy(), 0, atol=1e-12)


def test_q11_loss_matches_loops_and_masks():
    rs = np.random.RandomState(4)
    yt = np.zeros((2, 3, 48, 3)); yt[..., 0] = rs.rand(2, 3, 48) < 0.3
    yt[..., 1] = (rs.rand(2, 3, 48) < 0.5) * yt[..., 0]; yt[..., 2] = rs.rand(2, 3, 48) * yt[..., 0]
    yp = rs.rand(2, 3, 48, 3); yp[0, 0, 0, 0] = 0.0; yp[0, 0, 1, 0] = 1.0     # exercise the 1e-7 clip
    got = float(O.primary_loss(torch.tensor(yt), torch.tensor(yp)))
    assert abs(got - L.primary_loss_loops(yt, yp)) < 1e-12
    # nothing played: replay/volume terms are masked -> only 2*(-log(1-1e-7)) remains beyond bce_note
    yt0 = np.zeros((1, 1, 48, 3)); yp0 = np.full((1, 1, 48, 3), 0.25)
    want = -np.log(0.75) - np.log1p(-1e-7)
    assert abs(float(O.primary_loss(torch.tensor(yt0), torch.tensor(yp0))) - want) < 1e-9


def test_a16_nadam_matches_scalar_loop():
    g_seq = [0.3, -0.1, 0.25]
    p = {"w": torch.tensor([1.5], dtype=torch.float64)}
    st = O.NadamState()
    for g in g_seq:
        p = O.nadam_step(p, {"w": torch.tensor([g], dtype=torch.float64)}, st)
    assert abs(float(p["w"]) - L.nadam_scalar_loop(1.5, g_seq)) < 1e-14
    assert st.iterations == 3


def test_dropout_is_inverted_scaling():
    x = torch.ones(4, 5)
    m = torch.tensor(helpers.keep_mask(7, 4, 0.5, 4, 5))
    y = O.dropout(x, m, 0.5)
    assert set(np.unique(y.numpy())) <= {0.0, 2.0}
    big = helpers.keep_mask(7, 1, 0.2, 4096, 3)
    assert abs(big.mean() - 0.8) < 0.02
    assert abs(helpers.keep_mask(9, 6, 0.5, 4096, 256).mean() - 0.5) < 0.005


def test_generation_semantics():
    p = O.init_params(CFG, 0)
    sty = O.compute_genre(1)
    assert abs(sty.sum() - 1) < 1e-12 and np.count_nonzero(sty) == 6 and sty[3] == 1 / 6
    steps = 2
    u = np.random.RandomState(5).random_sample(2 * 48 * steps)
    ev_i, info_i = O.generate(p, CFG, [sty], steps, u, mode="incremental")
    ev_l, info_l = O.generate(p, CFG, [sty], steps, u, mode="literal")
    assert np.array_equal(ev_i[..., :2], ev_l[..., :2])
    np.testing.assert_allclose(ev_i[..., 2], ev_l[..., 2], atol=1e-6)
    # Q8: one uniform per note plus one more per PLAYED note
    assert info_i["uniforms_used"] == 48 * steps + int(ev_i[..., 0].sum())
    # replay only where played; volume only where played
    assert np.all(ev_i[..., 1] <= ev_i[..., 0]) and np.all((ev_i[..., 2] != 0) <= (ev_i[..., 0] == 1))


def test_q7_silence_raises_temperature_from_first_step():
    g = O.Generation(CFG, O.compute_genre(0))
    assert g.silent_time == 16 and g.beat_memory.sum() == 0
    g.end_time(0)                      # silent step: 17 >= 16 -> +0.1 immediately
    assert abs(g.temperature - 1.1) < 1e-12 and g.beat_memory[-1, 0] == 1
    g.next_note[3, 0] = 1
    g.end_time(1)
    assert g.temperature == 1 and g.silent_time == 0 and g.beat_memory[-1, 1] == 1


def test_q8_q9_choose_uses_le_and_float32_temperature():
    g = O.Generation(CFG, O.compute_genre(0))
    prob = np.zeros((48, 3), dtype=np.float32); prob[0] = (0.5, 0.25, 0.7)
    g.choose(prob, 0, O.UniformStream(np.array([0.5, 0.25])))          # u == p counts as a hit
    assert g.next_note[0].tolist() == [1, 1, np.float32(0.7)]
    t = O.apply_temperature(np.array([0.0, 1.0, 0.5], dtype=np.float32), 1.1)
    assert t.dtype == np.float32 and t[0] == 0 and t[1] == 1 and abs(t[2] - 0.5) < 1e-7


def test_golden_fixture_is_reproduced():
    z = np.load(GOLD)
    p = O.init_params(CFG, 0, torch.float64)
    notes, chosen, beat, style, target = O.synthetic_batch(CFG, 2, 4, 1234, torch.float64)
    probs = O.model_forward(p, CFG, notes, chosen, beat, style)
    np.testing.assert_allclose(probs.numpy(), z["predict_probs"], atol=1e-12)
    masks = helpers.oracle_masks(CFG, 2, 4, 7)
    loss, _, grads = O.loss_and_grads(p, CFG, notes, chosen, beat, style, target, masks)
    assert abs(float(loss) - float(z["train_loss"])) < 1e-12
    for k, g in grads.items():
        np.testing.assert_allclose(g.numpy().ravel()[:16], z[f"grad_head/{k}"], atol=1e-13)


def test_fp32_oracle_close_to_fp64():
    p64 = O.init_params(CFG, 0, torch.float64)
    p32 = {k: v.float() for k, v in p64.items()}
    b64 = O.synthetic_batch(CFG, 2, 8, 1234, torch.float64)
    b32 = [t.float() for t in b64]
    a = O.model_forward(p64, CFG, *b64[:4]); b = O.model_forward(p32, CFG, *b32[:4])
    assert helpers.rel_err(b.numpy(), a.numpy()) < 1e-5


def test_visualize_labels_table_matches_reference_layout():
    """visualize.py:27-40: header + one (genre, artist path) row per style, genre-major."""
    import visualize
    import constants
    t = visualize.style_labels()
    assert t.shape == (1 + constants.NUM_STYLES, 2) and list(t[0]) == ['Genre', 'Artist']
    assert t[1][0] == constants.genre[0] and t[1][1] == constants.styles[0][0]
    assert t[-1][1] == constants.styles[-1][-1]
