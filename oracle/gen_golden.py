"""TEST INFRASTRUCTURE ONLY — generate tests/golden/* from the LIVE reference.

Run in the build container (needs /root/reference):  python oracle/gen_golden.py
Everything written here is an OUTPUT of the unmodified reference code
(idelucs.kmers.kmer_counts, idelucs.utils.kmersFasta / AugmentFasta / transforms,
idelucs.LossFunctions.IID_loss) on the bundled Example/*.fas files or on
deterministic integer-generated inputs; the two FASTA files are copied (gzip) so the GPU
box, which has no /root/reference, can run the same inputs.
"""
import gzip
import hashlib
import json
import os
import random
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)

import ref_live  # noqa: E402
import idelucs_oracle as orc  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def loss_inputs(B, C, salt):
    """Deterministic softmax-like rows from integer arithmetic only (portable)."""
    b = np.arange(B, dtype=np.int64)[:, None]
    c = np.arange(C, dtype=np.int64)[None, :]
    h = (b * 2654435761 + c * 40503 + salt * 97 + (b * c) * 7919) % 1000003
    w = (h % 1000 + 1).astype(np.float32)
    # sharpen some rows so that joint entries span many orders of magnitude
    w = np.where((b + salt) % 3 == 0, w * w * w, w).astype(np.float32)
    return (w / w.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)


def main():
    ref = ref_live.load()
    import torch
    os.makedirs(GOLD, exist_ok=True)
    meta = {"reference_version": list(ref.__version__), "numpy": np.__version__}
    import sklearn
    meta["sklearn"] = sklearn.__version__

    # ---- 1. tiny known-answer tests for kmer_counts (kmers.pyx:2-50) ----
    kats = []
    cases = [(b"ACGTNACGTACGTTTT", 2), (b"ANCG", 1), (b"", 4), (b"ACG", 4), (b"ACGT", 4),
             (b"acgtACGTAC", 3), (b"AAAAAAAAAAAA", 6), (b"ACGTACGTNNACGTACGTAC-GTACGTTTGACA", 4),
             (b"ACGTACGTAGCTAGCTAGCTAGGGATCCCTA" * 9, 5), (bytes(range(256)) + b"ACGTACGTACGT", 6)]
    rng = np.random.default_rng(12345)
    for L in (1, 5, 6, 7, 63, 64, 65, 127, 128, 129, 1000, 4097):
        s = rng.choice(np.frombuffer(b"ACGTACGTACGTACGTN", dtype=np.uint8), size=L).tobytes()
        cases += [(s, 4), (s, 5), (s, 6)]
    for s, k in cases:
        c = np.zeros(4 ** k, dtype=np.int32)
        ref.kmer_counts(bytearray(s), k, c)
        nz = np.nonzero(c)[0]
        kats.append({"seq_hex": s.hex(), "k": k, "nz_idx": nz.tolist(), "nz_val": c[nz].tolist()})
    with open(os.path.join(GOLD, "kmer_kats.json"), "w") as fh:
        json.dump(kats, fh)

    # ---- 2. bundled FASTA files: counts / frequencies / x_train from the reference ----
    files = {}
    for stem in ("Influenza-A", "Actinopterygii"):
        src = os.path.join(ref_live.REFERENCE_ROOT, "Example", stem + ".fas")
        dst = os.path.join(GOLD, stem + ".fas.gz")
        with open(src, "rb") as fi, gzip.GzipFile(dst, "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
        gt = os.path.join(ref_live.REFERENCE_ROOT, "Example", stem + "_GT.tsv")
        shutil.copyfile(gt, os.path.join(GOLD, stem + "_GT.tsv"))
        entry = {}
        recs = orc.read_fasta(src)
        for k in (4, 5, 6):
            names, freq = ref.kmersFasta(src, k=k)
            counts = np.zeros((len(recs), 4 ** k), dtype=np.int32)
            for i, (_, seq) in enumerate(recs):
                ref.kmer_counts(seq, k, counts[i])
            entry[f"k{k}"] = {"n": len(names), "counts_sha256": sha(counts), "counts_sum": int(counts.sum()),
                              "counts_row0": counts[0, :8].tolist(), "freq64_sha256": sha(freq),
                              "freq32_sha256": sha(freq.astype(np.float32)),
                              "names_sha256": hashlib.sha256("\n".join(names).encode()).hexdigest()}
        files[stem] = entry

    # ---- 3. AugmentFasta with exported mutation lists ----
    src = os.path.join(ref_live.REFERENCE_ROOT, "Example", "Influenza-A.fas")
    aug = {}
    for n_mimics in (3,):
        # (a) the reference's own output, seeds as models.py:18-21 sets them
        for k in (4, 5, 6):
            np.random.seed(0)
            random.seed(0)
            x_train = ref.utils.AugmentFasta(src, n_mimics, k=k)
            sel = np.arange(0, x_train.shape[0], 211)
            aug[f"k{k}"] = {"shape": list(x_train.shape), "x_train_sha256": sha(x_train)}
            np.savez_compressed(os.path.join(GOLD, f"influenza_xtrain_rows_k{k}.npz"),
                                rows=sel, x=x_train[sel])
        # (b) the same RNG stream re-run with recording wrappers -> mutation lists
        np.random.seed(0)
        random.seed(0)
        U = ref.utils
        tfs = [U.transition_transversion(1e-2, 0.5e-2), U.transition(1e-2), U.transversion(0.5e-2)]
        tfs += [U.Random_N(20) for _ in range(n_mimics - 2)]
        pos_all, byte_all, offs = [], [], [0]
        for tf in tfs:
            rec = orc.RecordingTransform(tf)
            ref.kmersFasta(src, k=4, transform=rec)
            for pos, nb in rec.edits:
                pos_all.append(pos.astype(np.int32))
                byte_all.append(nb)
                offs.append(offs[-1] + len(pos))
        np.savez_compressed(os.path.join(GOLD, "influenza_edits_seed0.npz"),
                            pos=np.concatenate(pos_all), newbyte=np.concatenate(byte_all),
                            offsets=np.asarray(offs, dtype=np.int64), n_passes=len(tfs))
        aug["edits_per_pass"] = [int(sum(len(p) for p in pos_all[i * 949:(i + 1) * 949])) for i in range(len(tfs))]
    files["Influenza-A"]["augment_seed0_nmimics3"] = aug

    # inference featuriser (utils.py:400-405) float64 standardised profiles
    for stem in ("Influenza-A", "Actinopterygii"):
        srcf = os.path.join(ref_live.REFERENCE_ROOT, "Example", stem + ".fas")
        ds = ref.utils.SequenceDataset(srcf, k=6)
        files[stem]["inference_k6"] = {"kmers_sha256": sha(ds.kmers), "kmers32_sha256": sha(ds.kmers.astype(np.float32)),
                                       "row0": ds.kmers[0, :6].tolist()}
        np.savez_compressed(os.path.join(GOLD, f"{stem}_inference_rows_k6.npz"),
                            rows=np.arange(0, ds.kmers.shape[0], 97), x=ds.kmers[::97])

    # ---- 4. IID_loss / compute_joint forward + autograd gradients ----
    loss = []
    for (B, C, lamb, salt) in [(512, 5, 2.8, 1), (332, 5, 2.8, 2), (287, 5, 1.0, 3), (256, 3, 2.8, 4),
                               (512, 12, 2.8, 5), (512, 200, 2.8, 6), (64, 200, 2.0, 7), (1, 5, 2.8, 8),
                               (1024, 33, 2.5, 9), (7, 2, 2.8, 10)]:
        z1 = torch.from_numpy(loss_inputs(B, C, salt)).requires_grad_(True)
        z2 = torch.from_numpy(loss_inputs(B, C, salt + 100)).requires_grad_(True)
        val = ref.LossFunctions.IID_loss(z1, z2, lamb=lamb)
        val.backward()
        joint = ref.LossFunctions.compute_joint(z1.detach(), z2.detach()).numpy()
        z1d, z2d = z1.detach().double().requires_grad_(True), z2.detach().double().requires_grad_(True)
        val64 = ref.LossFunctions.IID_loss(z1d, z2d, lamb=lamb)
        val64.backward()
        rows = np.unique(np.linspace(0, B - 1, min(B, 16)).astype(np.int64))
        loss.append({"B": B, "C": C, "lamb": lamb, "salt": salt, "loss32": float(val.item()), "loss64": float(val64.item()),
                     "rows": rows.tolist(),
                     "dz1_rows64": z1d.grad.numpy()[rows].tolist(), "dz2_rows64": z2d.grad.numpy()[rows].tolist(),
                     "dz1_rows32": z1.grad.numpy()[rows].astype(np.float64).tolist(),
                     "joint_diag32": np.diag(joint).astype(np.float64).tolist()[:16],
                     "z1_sha256": sha(z1.detach().numpy())})
    with open(os.path.join(GOLD, "iid_loss.json"), "w") as fh:
        json.dump(loss, fh)

    meta["files"] = files
    with open(os.path.join(GOLD, "golden.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
