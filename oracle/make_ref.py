"""TEST / BASELINE INFRASTRUCTURE — stages the UNMODIFIED reference for the GPU box.

``/root/reference`` does not exist on the GPU box, and the reference is Python, so there is nothing to compile:
this recipe copies the five modules of the window-preparation path (recordutil, waveform_noise, paramutil, pathutil,
timelog) and the 37 params.json byte for byte into the git-ignored ``oracle/_ref/`` (it travels with the gpurun
snapshot like a built ``.so``; it never enters the history).  ``bench.py --impl reference`` and the ``cpu_baseline`` leg
then time the reference itself through ``oracle/ref_harness.py``'s stub shim (``cpu_baseline.kind: "reference"``) and fall
back to the port (``oracle/ref_port.py``, kind "port") only when ``oracle/_ref`` is absent.  Nothing in the product
imports it.  Run by ``__graft_entry__.build()`` when ``/root/reference`` is present:   python oracle/make_ref.py
"""
import glob
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
MODULES = ('recordutil.py', 'waveform_noise.py', 'paramutil.py', 'pathutil.py', 'timelog.py')


def available():
  return all(os.path.isfile(os.path.join(REF_DIR, m)) for m in MODULES)


def make(reference_path='/root/reference'):
  """Copy the path's modules + configs; returns the manifest (sha256 per file) or None if the reference is absent."""
  if not os.path.isfile(os.path.join(reference_path, 'recordutil.py')):
    return None
  os.makedirs(REF_DIR, exist_ok=True)
  manifest = {}
  files = [os.path.join(reference_path, m) for m in MODULES] + sorted(glob.glob(os.path.join(reference_path, 'waveform_*', 'params.json')))
  for src in files:
    rel = os.path.relpath(src, reference_path)
    dst = os.path.join(REF_DIR, rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    with open(dst, 'rb') as f:
      manifest[rel] = hashlib.sha256(f.read()).hexdigest()
  with open(os.path.join(REF_DIR, 'MANIFEST.json'), 'w') as f:
    json.dump({'source': reference_path, 'files': manifest}, f, indent=1)
  return manifest


if __name__ == '__main__':
  m = make()
  print('oracle/_ref: %s' % ('%d files' % len(m) if m else 'reference not present, nothing staged'))
