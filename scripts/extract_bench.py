import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    r=d.get("roofline",{})
    print(json.dumps({"value":round(d["value"]/1e6,1),"ms":round(d["ms_per_step"],2),"frac":round(r.get("frac",0),4),"kernel_ms":round(r.get("kernel_ms_per_step",0),2),"phase":{k:round(v,2) for k,v in r.get("phase_ms_per_step",{}).items()},"checksum":d.get("checksum")}))
