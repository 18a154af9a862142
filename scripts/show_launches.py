import csv, json, sys
lines=[l for l in open('/root/repo/gpurun_out/launches.csv') if not l.startswith('==')]
seq=[(x['Kernel Name'].split('(')[0][:28], float(x['Metric Value'].replace(',',''))/1000) for x in csv.DictReader(lines)]
steps=[]; cur=None
for n,t in seq:
    if 'k_codebook_query' in n:
        if cur: steps.append(cur)
        cur=[]
    if cur is not None: cur.append((n,t))
if cur: steps.append(cur)
for i in (0,1,5,10,20,30):
    if i < len(steps):
        print(i, [(n.replace('void ',''), round(t,1)) for n,t in steps[i] if 'step' in n or 'cosine' in n or 'query' in n][:8], round(sum(t for n,t in steps[i][:6])))
d=json.load(open('/root/repo/gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('engine_stats'))
