for w in 4 8 12 16 24; do echo "readers $w"; SCGRHC_READERS=$w timeout 300 python tools/dropin_probe.py 500 2>&1 | grep "eager 32\|streamed 32"; done
python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))"
