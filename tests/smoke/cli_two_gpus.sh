#!/bin/bash
# End-to-end smoke of the public CLIs on 2 GPUs (not a pytest: needs a 2-GPU box and ~2 minutes):
#   torchrun train_spnet.py --parallel (frozen phase + unfrozen phase, checkpoints, evaluate tail), then
#   torchrun predict_spnet.py (rank-sharded, merged CSV) against the single-process CSV.
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
for split, n, seed in (("Train", 64, 1000), ("Val", 16, 5000)):
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
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    "$ROOT/train_spnet.py" --parallel -d "$WORK" -b 16 -e 3 --freeze_fac 0.75 --frozen_epochs 1 --model_type big -n \
    --predict_path "$WORK/Val" > train.log 2>&1 || { grep -n -B2 -A25 "Traceback" train.log | head -80; exit 1; }
grep -E "Epoch|mAP|SPNet execution completed" train.log | tail -8
ls *.hdf5 *.h5 2>/dev/null | head
python "$ROOT/predict_spnet.py" -w final_weights.hdf5 -d "$WORK/Val" -b 4 --model_type big --no-png > p1.log 2>&1 || { tail -20 p1.log; exit 1; }
cp logs/Predicting/hawley_spnet.csv single.csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    "$ROOT/predict_spnet.py" -w final_weights.hdf5 -d "$WORK/Val" -b 4 --model_type big --no-png > p2.log 2>&1 || { tail -20 p2.log; exit 1; }
cmp single.csv logs/Predicting/hawley_spnet.csv && echo "CLI_SMOKE_OK: sharded predict CSV identical to single-process CSV ($(wc -l < single.csv) rows)"
