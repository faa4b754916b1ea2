# round 2, call T: per-launch times of the selection kernels at a mid level of config 4 (ncu, kernel filter)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:'k_select|k_pick|k_tau|k_sim_sample|k_sim_seljobs' -s 1400 -c 160 --csv --log-file gpurun_out/r02_sel_launches_cfg4_mid.csv python bench.py --config 4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_sel.log 2>&1
tail -c 200 gpurun_out/r02_ncu_sel.log
python - <<'PY'
import csv, io, re
txt=open('gpurun_out/r02_sel_launches_cfg4_mid.csv').read()
lines=[l for l in txt.splitlines() if not l.startswith('==')]
rows=list(csv.DictReader(io.StringIO("\n".join(lines))))
cur=None
for r in rows:
    name=re.sub(r'\(.*','',r['Kernel Name']).replace('void ','').replace('iq::','').replace('iqimpl::','')
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    if r['Metric Name']=='gpu__time_duration.sum':
        us = v/1e3 if u.startswith('n') else (v if u.startswith('u') else v*1e3)
        cur=[name,us,r['Grid Size'],0.0]
    elif r['Metric Name']=='dram__bytes_read.sum' and cur:
        mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
        cur[3]=v*mult/1e6
        if cur[1] > 6: print("%-24s %8.1f us %-14s %8.1f MB"%tuple(cur))
PY
