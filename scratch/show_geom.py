import json,csv,sys,collections
v=sys.argv[1]
d=json.load(open('gpurun_out/r02_kernels_isolated_%s.json'%v))
for k,val in d['kernels'].items(): print('%-50s %8.1f us %7.1f GB/s  %.3f'%(k, val['us'], val['achieved_gbs'], val['frac_of_measured_peak']))
rows=list(csv.reader(open('gpurun_out/r02_ncu_geom_%s.csv'%v)))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[hi]; ki,mi,vi=H.index('Kernel Name'),H.index('Metric Name'),H.index('Metric Value'); ii=H.index('ID')
per=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    per.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for k,val in list(per.items())[:3]+list(per.items())[6:7]:
    print(k, {a.split('.')[0].replace('smsp__','').replace('sm__','')[:22]:b for a,b in val.items()})
