"""Generate the committed fixtures under tests/golden/ from the read-only reference checkout.

Run in the build container (where /root/reference exists):  python tests/golden/make_fixtures.py
/root/reference does not exist on the GPU box, so tests, smoke() and bench.py only read what this
script wrote.  Nothing here is reference SOURCE code: the outputs are the reference's shipped data
(checkpoint tensors, xyz inputs, its own predictions = golden vectors).

Outputs
  checkpoints/{decay_model_weights,model_weights,model2_weights}.*   byte copies of models/* (inputs
                              to the tensor-bundle reader; reference infer.py:57)
  mixed.npz                   every data/mixed.tar.gz system, packed (names, offsets, xyz f32, Z, Q, labels)
  val871.npz                  models/model_systems/{val_names,test_pred_charges,test_lab_charges}.npy
                              (golden predictions of decay_model_weights at pad N=41)
  protein.npz                 protein/6qlp_capped.xyz + protein/preds.npy (golden, pad N=n=2220)
  xyz/                        a handful of raw xyz files (+ label npy, + one splits.npy) for the loader/CLI tests
"""
import io
import os
import shutil
import tarfile

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ZNUM = {'H': 1, 'C': 6, 'N': 7, 'O': 8, 'F': 9, 'P': 15, 'S': 16, 'Cl': 17, 'Br': 35}


def parse(text):
    lines = text.splitlines()
    Q = np.float32(lines[1].strip().split()[0])
    sym, xyz = [], []
    for line in lines[2:]:
        d = line.split()
        if not d:
            continue
        sym.append(d[0])
        xyz.append([d[1], d[2], d[3]])
    return sym, np.array(xyz, dtype=np.float32), Q


def main():
    ck = os.path.join(OUT, "checkpoints")
    os.makedirs(ck, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "models"))):
        if "_weights." in f:
            shutil.copyfile(os.path.join(REF, "models", f), os.path.join(ck, f))

    # ---- mixed set
    tf = tarfile.open(os.path.join(REF, "data", "mixed.tar.gz"))
    members = {m.name: m for m in tf.getmembers() if m.isfile()}
    xyz_names = sorted(n for n in members if n.endswith(".xyz"))
    names, offs, coords, Zs, Qs, labs, has_lab = [], [0], [], [], [], [], []
    raw_text = {}
    for n in xyz_names:
        text = tf.extractfile(members[n]).read().decode()
        sym, xyz, Q = parse(text)
        stem = os.path.basename(n)[:-4]
        raw_text[stem] = text
        names.append(stem)
        coords.append(xyz)
        Zs.append(np.array([ZNUM[s] for s in sym], dtype=np.int8))
        Qs.append(Q)
        offs.append(offs[-1] + len(sym))
        ln = n[:-4] + ".npy"
        if ln in members:
            y = np.load(io.BytesIO(tf.extractfile(members[ln]).read())).astype(np.float32).reshape(-1)
            has_lab.append(True)
        else:
            y = np.zeros(len(sym), np.float32)
            has_lab.append(False)
        labs.append(y)
    np.savez_compressed(os.path.join(OUT, "mixed.npz"), names=np.array(names), offsets=np.array(offs, np.int32),
                        xyz=np.concatenate(coords), Z=np.concatenate(Zs), Q=np.array(Qs, np.float32),
                        labels=np.concatenate(labs), has_labels=np.array(has_lab))

    # ---- 871 golden validation systems
    ms = os.path.join(REF, "models", "model_systems")
    vn = np.load(os.path.join(ms, "val_names.npy"), allow_pickle=True)
    pred = np.load(os.path.join(ms, "test_pred_charges.npy"))
    lab = np.load(os.path.join(ms, "test_lab_charges.npy"))
    np.savez_compressed(os.path.join(OUT, "val871.npz"), names=np.array([str(x) for x in vn]),
                        pred=pred.astype(np.float32), lab=lab.astype(np.float32))

    # ---- protein
    tp = tarfile.open(os.path.join(REF, "data", "protein.tar.gz"))
    text = tp.extractfile("protein/6qlp_capped.xyz").read().decode()
    sym, xyz, Q = parse(text)
    preds = np.load(io.BytesIO(tp.extractfile("protein/preds.npy").read()))
    np.savez_compressed(os.path.join(OUT, "protein.npz"), xyz=xyz, Z=np.array([ZNUM[s] for s in sym], np.int8),
                        Q=np.float32(Q), preds=preds.astype(np.float32))

    # ---- a few raw files for the loader / CLI tests
    xd = os.path.join(OUT, "xyz")
    os.makedirs(xd, exist_ok=True)
    vset = [str(x) for x in vn]
    picks = [n for n in vset if n.startswith("dsgdb9nsd")][:3] + [n for n in vset if n.startswith("SSI")][:2]
    charged = [n for n, q in zip(names, Qs) if q != 0 and n in set(vset)][:2]
    for stem in picks + charged:
        with open(os.path.join(xd, stem + ".xyz"), "w") as f:
            f.write(raw_text[stem])
        ln = "mixed/" + stem + ".npy"
        if ln in members:
            with open(os.path.join(xd, stem + ".npy"), "wb") as f:
                f.write(tf.extractfile(members[ln]).read())
        sp = "mixed/" + stem + "splits.npy"
        if sp in members:
            with open(os.path.join(xd, stem + "splits.npy"), "wb") as f:
                f.write(tf.extractfile(members[sp]).read())
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
