#!/bin/bash
# End-to-end smoke of the public CLIs on one GPU (not a pytest; ~2 minutes): train_spnet.py with the host-side
# AugmentOnTheFly callback, train_spnet.py --device_data (training set and augmentation on the GPU), evaluate_spnet.py,
# predict_spnet.py --stream.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
WORK=$(mktemp -d)
cd "$WORK"
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from PIL import Image
from spnet_b200 import fake_espi
from spnet import utils
import os
for split, n, seed in (("Train", 48, 1000), ("Val", 16, 5000)):
    os.makedirs(split, exist_ok=True)
    i = 0
    while i < n:
        img, rows = fake_espi.make_frame(seed); seed += 1
        try:
            utils.build_Y_from_rows([rows], pred_grid=[6, 6, 2])
        except AssertionError:
            continue
        Image.fromarray(img.reshape(img.shape[0], img.shape[1])).save("%s/steelpan_%07d.png" % (split, i))
        open("%s/steelpan_%07d.csv" % (split, i), "w").write("\n".join(",".join(str(v) for v in r) for r in rows))
        i += 1
print("dataset written")
PY
run() { name=$1; shift; "$@" > $name.log 2>&1 || { echo "FAILED: $name"; grep -n -B2 -A25 "Traceback" $name.log | head -60; exit 1; }; }
run train_aug python "$ROOT/train_spnet.py" -d "$WORK" -b 16 -e 2 --model_type big --predict_path "$WORK/Val"
grep -E "Epoch|execution completed" train_aug.log | tail -3
run train_dev python "$ROOT/train_spnet.py" -d "$WORK" -b 16 -e 2 --model_type big --device_data --predict_path "$WORK/Val" -w weights_dev.hdf5
grep -E "Epoch|execution completed" train_dev.log | tail -3
run evaluate python "$ROOT/evaluate_spnet.py" -w final_weights.hdf5 -d "$WORK/Val/" -b 4 --model_type big
tail -3 evaluate.log
run predict_stream python "$ROOT/predict_spnet.py" -w final_weights.hdf5 -d "$WORK/Val" -b 4 --model_type big --no-png --stream 8
run predict_whole python "$ROOT/predict_spnet.py" -w final_weights.hdf5 -d "$WORK/Val" -b 4 --model_type big --no-png -l logs/Whole/
cmp logs/Predicting/hawley_spnet.csv logs/Whole/hawley_spnet.csv && echo "CLI_SMOKE_OK ($(wc -l < logs/Whole/hawley_spnet.csv) CSV rows)"
