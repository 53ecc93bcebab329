"""Plain pinned-host -> device copy bandwidth (the ceiling of the end-to-end leg) for several chunk sizes — GPU box."""
import json, sys, torch
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 9.6
n = int(gb * 1e9) // 8
host = torch.empty(n, dtype=torch.float64, pin_memory=True); host.fill_(1.0)
dev = torch.empty(n, dtype=torch.float64, device='cuda')
def timed(chunks, streams=1):
  ss = [torch.cuda.Stream() for _ in range(streams)]
  step = -(-n // chunks)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for s in ss: s.wait_stream(torch.cuda.current_stream())
  for k in range(chunks):
    with torch.cuda.stream(ss[k % streams]):
      dev[k * step:(k + 1) * step].copy_(host[k * step:(k + 1) * step], non_blocking=True)
  for s in ss: torch.cuda.current_stream().wait_stream(s)
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b)
for chunks, streams in ((1, 1), (20, 1), (20, 2), (100, 1), (100, 4)):
  timed(chunks, streams)
  ms = min(timed(chunks, streams) for _ in range(3))
  print(json.dumps(dict(gb=gb, chunks=chunks, streams=streams, ms=round(ms, 2), gbs=round(gb / ms * 1e3, 2))))
