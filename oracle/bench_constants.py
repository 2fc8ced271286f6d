"""Record the first-step loss of bench.py's two workloads from the UNMODIFIED reference, for bench.py to assert.

Run in the build container (needs /root/reference):   python oracle/bench_constants.py
Writes tests/golden/bench_first_step.json. Test infrastructure: bench.py only reads the JSON constant.

* rec: reference RecognitionModel (seed 1234) + torch.nn.CTCLoss on bench.make_batch("rec", rank 0), 64x(64x800),
  computed in fp64 by running conv / gru / output in sequence (models.py:266 hard-codes .float()) and in fp32.
* det: reference DetectionModel (seed 1234) + balanced_cross_entropy_loss on bench.make_batch("det", rank 0),
  32x(1024x1024), forward + loss only under no_grad (the autograd graph of this batch needs ~65 GB), fp32 and fp64.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import bench  # noqa: E402
from make_golden import import_reference  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    models, train_detection, train_rec, alphabet = import_reference()
    out = {"seed": 1234, "rank": 0}
    b = bench.make_batch("rec", 0, "cpu")
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        torch.manual_seed(1234)
        m = models.RecognitionModel(alphabet).train().to(dt)
        with torch.no_grad():
            f = m.conv(b["image"].to(dt))
            f = torch.permute(f, (3, 0, 1, 2)).reshape(f.shape[3], f.shape[0], -1)
            f, _ = m.gru(f)
            lp = m.output(f)
            loss = torch.nn.CTCLoss()(lp, b["targets"], b["input_lengths"], b["target_lengths"])
        out[f"rec_loss_{name}"] = float(loss)
        print("rec", name, float(loss), flush=True)
    b = bench.make_batch("det", 0, "cpu")
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        torch.manual_seed(1234)
        m = models.DetectionModel().train().to(dt)
        with torch.no_grad():
            y = m(b["image"].to(dt))
            loss = train_detection.balanced_cross_entropy_loss(y, b["mask"].to(dt))
        out[f"det_loss_{name}"] = float(loss)
        out[f"det_prob_mean_{name}"] = float(y.mean())
        del y
        print("det", name, float(loss), flush=True)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bench_first_step.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
