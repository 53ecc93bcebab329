"""scgrhc — B200-native engine behind the drop-in recordutil / waveform_noise modules."""
from . import _native  # noqa: F401
from .engine import (Plan, WindowStore, plan_cohort, plan_record, plan_uniform, prepare_windows, prepare_subsets,  # noqa: F401
                     resolve_columns, shard_records, allreduce_minmax, event_table)
