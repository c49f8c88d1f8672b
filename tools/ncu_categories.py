import sys,csv,collections,re
sys.path.insert(0,'/root/repo/tools')
from ncu_by_line import line_table
dis, csvp, mangled = sys.argv[1:4]
table=line_table(dis,mangled)
rows=list(csv.reader(open(csvp)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[hi]
ci={n:hdr.index(n) for n in ("Address","# Samples","Instructions Executed","Thread Instructions Executed")}
def cat(loc):
    f,l=loc
    if f=='rt_split.cuh':
        if l<=476: return 'handout/prefetch pipeline'
        if l<=512: return 'ray setup (set xform, plane)'
        if l<=572: return 'walk control + interior box'
        if l<=634: return 'mesh entry + suspend'
        if l<=680: return 'analytic leaf'
        return 'store/next chunk'
    if f=='rt_device.cuh':
        if l<=210: return 'vector math (div/normalize/minmax)'
        if l<=250: return 'xform bracket+quat_rotate'
        if l<=315: return 'xform_eval'
        if l<=375: return 'to_local (rotate_exact, scale)'
        if l<=420: return 'box test'
        if l<=465: return 'tri'
        if l<=510: return 'sphere'
        if l<=525: return 'plane'
        return 'rect'
    if f=='rt_trace.cuh':
        if l<=50: return 'local_ray_finish (3 rcp)'
        if l<=96: return 'load node/shape'
        return 'shape_xform'
    if f=='rt_render.cuh': return 'IO (queue at/decode/store/bq_push)'
    if 'atomic' in f: return 'atomics'
    if 'intrinsics' in f or 'pipeline' in f: return 'intrinsics (ldg/shfl/ballot/cp.async)'
    return f
agg=collections.defaultdict(lambda: [0,0,0])
base=None
for r in rows[hi+1:]:
    if len(r)<len(hdr)-1 or r[0]=="Address": continue
    a=int(r[0],16)
    if base is None: base=a
    loc=table.get(a-base,("?",0))
    c=cat(loc)
    agg[c][0]+=float(r[ci["# Samples"]] or 0); agg[c][1]+=float(r[ci["Instructions Executed"]] or 0); agg[c][2]+=float(r[ci["Thread Instructions Executed"]] or 0)
ts=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
for c,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print("%-42s instr %5.1f%%  samples %5.1f%%  lanes %5.1f"%(c,100*v[1]/ti,100*v[0]/ts,v[2]/max(v[1],1)))
