"""Drop-in for the reference's ``paramutil`` (paramutil.py:5-33): JSON -> attributes.

Same class, same attribute names, same ``Params(path)`` call.  Differences, all opt-in or additive:
  * the five legacy configs the reference's own loader rejects with ``KeyError`` (waveform_01..05,
    SURVEY.md §0) load with documented defaults (``LEGACY_DEFAULTS``); ``Params(path, strict=True)``
    restores the reference's hard ``KeyError``;
  * optional keys read with ``.data.get`` only (absent from all 37 shipped params.json, so shipped
    behaviour is unchanged): ``split_seed`` (reproducible train/valid/test split), ``segment_stride``
    (seconds between window starts; default = ``segment_size``, i.e. the reference's non-overlapping windows), ``noise_std`` / ``noise_seed`` (train-time
    noise injection on SCG batches), ``normalisation`` ('zscore': per-window mean/std instead of min-max), ``bandpass`` / ``bandpass_order`` / ``bandpass_sos`` (zero-phase IIR filtering of the
    SCG channels), ``resample_rate`` (model sampling rate).  With none of them present the path is the reference's.
"""
import json
import os

LEGACY_DEFAULTS = {
  'chamber': '*',                    # waveform_01 predates chamber segmentation (project_log.txt:1-6): every chamber interval
  'checkpoint_dir_path': 'checkpoints',
  'comparison_dir_path': 'comparisons',
  'pred_top_dir_path': 'pred_top',
  'pred_rand_dir_path': 'pred_rand',
  'min_RHC': float('-inf'),          # no pressure floor before waveform_06
  'use_global_min_max': False,
}


class Params:
  def __init__(self, path, strict=False):
    self.path = path
    self.data = self.init_json(path)
    self.strict = strict
    g = self._get
    self.in_channels = g('in_channels')
    self.chamber = g('chamber')
    self.segment_size = g('segment_size')
    self.batch_size = g('batch_size')
    self.dir_path = g('dir_path')
    self.train_path = os.path.join(self.dir_path, g('train_path'))
    self.valid_path = os.path.join(self.dir_path, g('valid_path'))
    self.test_path = os.path.join(self.dir_path, g('test_path'))
    self.checkpoint_dir_path = os.path.join(self.dir_path, g('checkpoint_dir_path'))
    self.comparison_dir_path = os.path.join(self.dir_path, g('comparison_dir_path'))
    self.pred_top_dir_path = os.path.join(self.dir_path, g('pred_top_dir_path'))
    self.pred_rand_dir_path = os.path.join(self.dir_path, g('pred_rand_dir_path'))
    self.alpha = g('alpha')
    self.beta1 = g('beta1')
    self.beta2 = g('beta2')
    self.n_critic = g('n_critic')
    self.lambda_gp = g('lambda_gp')
    self.lambda_aux = g('lambda_aux')
    self.total_epochs = g('total_epochs')
    self.min_RHC = g('min_RHC')
    self.use_global_min_max = g('use_global_min_max')
    # additive, optional
    self.split_seed = self.data.get('split_seed')
    self.segment_stride = self.data.get('segment_stride')
    self.noise_std = self.data.get('noise_std')      # train-time Gaussian noise on SCG batches (extension, default off)
    self.noise_seed = self.data.get('noise_seed')
    self.bandpass = self.data.get('bandpass')              # [low_hz, high_hz]: zero-phase Butterworth on the SCG channels
    self.bandpass_order = self.data.get('bandpass_order')
    self.bandpass_sos = self.data.get('bandpass_sos')      # or explicit second-order sections
    self.bandpass_mode = self.data.get('bandpass_mode')    # 'exact' (default, bit-identical to scipy) | 'scan' (time-parallel)
    self.resample_rate = self.data.get('resample_rate')    # model sampling rate in Hz (native: 500)
    self.resample_mode = self.data.get('resample_mode')    # 'exact' (default, bit-identical to scipy) | 'fused' (one FMA per tap)
    self.train_layout = self.data.get('train_layout')      # multi-GPU jobs: 'gathered' (rank 0 holds the train loader) | 'sharded'
    self.normalisation = self.data.get('normalisation')    # 'minmax' (default = the reference) | 'zscore' (per-window mean/std)

  def _get(self, key):
    if key in self.data or self.strict or key not in LEGACY_DEFAULTS:
      return self.data[key]            # KeyError for a missing key, as the reference (paramutil.py:9-29)
    if key in ('chamber', 'min_RHC', 'use_global_min_max'):      # the keys that change which windows are produced
      import warnings
      warnings.warn('%s has no %r: using the documented legacy default %r (the reference raises KeyError here; '
                    'Params(path, strict=True) does too)' % (self.path, key, LEGACY_DEFAULTS[key]), stacklevel=3)
    return LEGACY_DEFAULTS[key]

  def init_json(self, path):
    with open(path, 'r') as f:
      return json.load(f)
