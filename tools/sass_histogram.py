"""Mnemonic histogram of the built library's SASS (cuobjdump -sass), whole library and the headline kernel alone."""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'scg-rhc-waveform_b200', 'scgrhc', 'libscgrhc.so')
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
head = 'window_planar_kernelILi3ELi128ELi3EfLi750ELb1'
tot, per, cur, nfun = collections.Counter(), collections.Counter(), None, 0
for line in txt.split('\n'):
  m = re.match(r'\s*Function : (\S+)', line)
  if m:
    cur, nfun = m.group(1), nfun + 1
    continue
  m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)', line)
  if m:
    tot[m.group(1)] += 1
    if cur and head in cur:
      per[m.group(1)] += 1
out = {'total_instructions': sum(tot.values()), 'functions': nfun, 'mnemonics': dict(tot.most_common()),
       'headline_kernel': {'name': 'scgrhc::window_planar_kernel<3,128,3,float,750,PLAIN>', 'instructions': sum(per.values()), 'mnemonics': dict(per.most_common())}}
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_histogram.json'), 'w'), indent=0)
print(out['total_instructions'], out['functions'], out['headline_kernel']['instructions'], list(per.most_common(12)))
