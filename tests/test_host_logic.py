"""CPU tests of host-side logic that needs neither a GPU nor the reference mount."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _ref_style_info_nce(z1, z2, temperature):
    """the reference's mask-gather formulation (idelucs/LossFunctions.py:65-98), restated"""
    n = z1.shape[0]
    feats = torch.nn.functional.normalize(torch.cat((z1, z2), 0).float(), dim=1)
    lab = torch.cat([torch.arange(n), torch.arange(n)])
    lab = (lab.unsqueeze(0) == lab.unsqueeze(1))
    sim = feats @ feats.T
    eye = torch.eye(2 * n, dtype=torch.bool)
    lab, sim = lab[~eye].view(2 * n, -1), sim[~eye].view(2 * n, -1)
    logits = torch.cat([sim[lab].view(2 * n, -1), sim[~lab].view(2 * n, -1)], dim=1) / temperature
    return torch.nn.functional.cross_entropy(logits, torch.zeros(2 * n, dtype=torch.long))


def test_info_nce_equals_reference_formulation():
    from idelucs_b200.LossFunctions import info_nce_loss
    torch.manual_seed(0)
    for n, d in ((5, 8), (64, 64), (257, 64)):
        a, b = torch.randn(n, d, requires_grad=True), torch.randn(n, d, requires_grad=True)
        want = _ref_style_info_nce(a, b, 0.85)
        got = info_nce_loss(a, b, 0.85)
        assert abs(want.item() - got.item()) < 1e-6
        gw = torch.autograd.grad(want, a)[0]
        gg = torch.autograd.grad(got, a)[0]
        assert torch.allclose(gw, gg, atol=1e-6)
        from idelucs_b200.LossFunctions import info_nce_loss_stacked
        st = info_nce_loss_stacked(torch.cat((a, b), 0), 0.85)      # the training step's stacked form
        assert abs(want.item() - st.item()) < 1e-6
        assert torch.allclose(gw, torch.autograd.grad(st, a)[0], atol=1e-6)


def test_check_sequence_shim_matches_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import idelucs_oracle as orc
    from idelucs_b200.utils import check_sequence
    for s in (b"acgtuUswkmyrbdhvnSWKMYRBDHV-ACGTN", b"AC GT\tAC\nGT\r", b"", b"NNNN"):
        assert bytes(check_sequence("h", bytearray(s))) == bytes(orc.check_sequence("h", bytearray(s)))
    for s in (b"ACGTX", b"ACxGT", b"AC\xe9"):
        with pytest.raises(ValueError) as e1:
            orc.check_sequence("hdr", bytearray(s))
        with pytest.raises(ValueError) as e2:
            check_sequence("hdr", bytearray(s))
        assert str(e1.value) == str(e2.value)
    for h in (">x", "#x", " x", "a\tb"):
        with pytest.raises(ValueError):
            check_sequence(h, bytearray(b"ACGT"))


def test_mimic_schedule_matches_reference_pass_order():
    from idelucs_b200 import featurise as ft
    v = ft.mimic_schedule(50)
    assert len(v) == 51 and [x.kind for x in v[:3]] == [ft.KIND_BOTH, ft.KIND_TRANSITION, ft.KIND_TRANSVERSION]
    assert all(x.kind == ft.KIND_RANDOM_N and x.n_bp == 20 for x in v[3:])
    assert (v[0].p1, v[0].p2, v[1].p1, v[2].p2) == (1e-2, 0.5e-2, 1e-2, 0.5e-2)
    assert [x.rng_id for x in v] == list(range(51))
    assert len(ft.mimic_schedule(1)) == 3          # the reference always runs passes 0..2 (utils.py:330-344)


def test_fasta_reader_quirks(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import idelucs_oracle as orc
    from idelucs_b200.seqset import read_fasta_raw
    f = tmp_path / "x.fas"
    f.write_bytes(b"#comment\n>s1 desc\nACGT\nacgt \n#mid\n>s2\n\nNNAC\nGT\n>s3\n")
    names, seqs = read_fasta_raw(str(f))
    want = orc.read_fasta(str(f))
    assert names == [n for n, _ in want] == ["s1 desc", "s2", "s3"]
    assert [bytes(orc.check_sequence(n, bytearray(s))) for n, s in zip(names, seqs)] == [bytes(s) for _, s in want]
    e = tmp_path / "empty.fas"
    e.write_bytes(b"")
    assert read_fasta_raw(str(e)) == ([""], [b""])
