"""Run on the GPU box: prepare a tiny cohort with the drop-in recordutil and leave the three pickled loaders in
gpurun_out/loader_fixture/ — they are committed under tests/golden/loaders/ and fed to the reference's UNMODIFIED
waveform_train / waveform_test / waveform_checkpoint in tests/test_reference_consumers.py (CPU, build container)."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'scg-rhc-waveform_b200'))
import numpy as np  # noqa: E402
from oracle import synth_ref  # noqa: E402
import recordutil  # noqa: E402
from scgrhc import wfdbio  # noqa: E402

out = os.path.join(ROOT, 'gpurun_out', 'loader_fixture')
data = os.path.join(out, 'data')
os.makedirs(data, exist_ok=True)
sig = synth_ref.DEFAULT_SIG_NAMES
meta = synth_ref.record_meta(60, events={'RA_1': 0, 'PA_1': 6})
for r in range(2):
  p = synth_ref.gen_record(0x5C6, 60 + r, 30000, kinds=synth_ref.kinds_for(sig), defect_scale=4)
  wfdbio.wrsamp('rec%d' % r, 500, ['g', 'g', 'g', 'mmHg'], sig, p, write_dir=data)
  with open(os.path.join(data, 'rec%d.json' % r), 'w') as f:
    json.dump(meta, f)
recordutil.PROCESSED_DATA_PATH = data
recordutil.wfdb = wfdbio
for f in ('loader_train.pickle', 'loader_valid.pickle', 'loader_test.pickle'):
  if os.path.exists(os.path.join(out, f)):
    os.remove(os.path.join(out, f))
params = types.SimpleNamespace(dir_path=out, train_path=os.path.join(out, 'loader_train.pickle'),
                               valid_path=os.path.join(out, 'loader_valid.pickle'), test_path=os.path.join(out, 'loader_test.pickle'),
                               in_channels=sig[:3], chamber='PA', segment_size=1.5, batch_size=16, min_RHC=-50,
                               use_global_min_max=False, split_seed=1)
recordutil.run(params)
print(open(os.path.join(out, 'record_log.txt')).read())
