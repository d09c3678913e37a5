"""Regenerates tests/golden/ from the reference checkout (run in the build container only).

1. Copies the reference's four test fixtures (core-test/src/test/resources/*.data: test INPUT
   data, not source code) gzip-compressed, so GPU-box tests never read /root/reference.
2. Runs the fp64 oracle over them and stores the known answers (accuracies, learned
   parameters) in known_answers.json.  The accuracies are what pins the oracle to the
   reference: README.md:9-15 publishes 65 / 42 / 76 / 96 / 99 % and
   DynamicNaiveBayesianClassifierTest.scala:13,19,25,33 asserts the thresholds.

    python tests/golden/make_golden.py
"""
import gzip
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/core-test/src/test/resources"
FIXTURES = ["robot_no_momentum", "robot_no_momentum_continuous", "robot_no_momentum_bivariate", "gaussian_mixture"]


def main():
    from dnbc_scala_b200 import dataset
    from oracle import oracle as orc

    for name in FIXTURES:
        with open(os.path.join(REF, name + ".data"), "rb") as src, \
                gzip.GzipFile(os.path.join(HERE, name + ".data.gz"), "wb", mtime=0) as dst:
            shutil.copyfileobj(src, dst)
    answers = {}
    for name in FIXTURES:
        ds = dataset.load(os.path.join(HERE, name + ".data.gz"))
        learn, test = orc.Data.of(ds.learn), orc.Data.of(ds.test)
        entry = {"N": ds.N, "V": ds.V, "learn_S": learn.S, "test_S": test.S, "runs": {}}
        hint_sets = [[1]] if name != "gaussian_mixture" else [[1], [2]]
        for quirk in (1, 0):
            for hints in hint_sets:
                h = hints if learn.C else []
                cen = None
                if h and h[0] == 2:  # deterministic seeds: extreme points of each edge
                    cen = {(0, j): [float(learn.cont[0, learn.labels == j].min()),
                                    float(learn.cont[0, learn.labels == j].max())] for j in range(ds.N)}
                m = orc.mle(learn, ds.N, ds.V, h or None, flags=orc.QUIRK_TRANSITIONS if quirk else 0,
                            init_centroids=cen)
                path, score, tie = orc.decode(m, test, orc.DECODE_REFERENCE)
                bpath, bscore, _ = orc.decode(m, test, orc.DECODE_BACKTRACE)
                entry["runs"][f"quirk{quirk}_k{h[0] if h else 0}"] = {
                    "accuracy_reference": orc.accuracy(test, path),
                    "accuracy_backtrace": orc.accuracy(test, bpath),
                    "score_sum": float(np.sum(score[np.isfinite(score)])),
                    "ties": int((tie != 0).sum()),
                    "pi": m.pi.tolist(), "A_rowsum": m.A.sum(1).tolist(),
                    "w": m.w.tolist(), "mu": m.mu.tolist(), "var": m.var.tolist(),
                    "init_centroids": None if cen is None else {f"{k[0]},{k[1]}": v for k, v in cen.items()},
                }
        answers[name] = entry
        for k, r in entry["runs"].items():
            print(name, k, round(r["accuracy_reference"], 3), round(r["accuracy_backtrace"], 3), "ties", r["ties"])
    with open(os.path.join(HERE, "known_answers.json"), "w") as fh:
        json.dump(answers, fh, indent=1)


if __name__ == "__main__":
    main()
