#!/usr/bin/env python
"""SASS instruction count per CUDA source line of one kernel (nvdisasm -g -c output)."""
import collections
import re
import sys

txt = open(sys.argv[1]).read()
name = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sec = txt.split('\n.text.%s:' % name)[1].split('\n.text.')[0]
cnt, cur = collections.Counter(), None
for line in sec.split('\n'):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        if cur is None or 'inlined at' not in line:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', line):
        cnt[cur] += 1
print('total', sum(cnt.values()))
for k, v in cnt.most_common(top):
    print('%s:%d %d' % (k[0], k[1], v))
