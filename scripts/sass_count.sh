#!/bin/bash
# Instruction count of the pose loops (K=4 and K=5 row groups) of the f16d scoring and fwd+bwd kernels.
O=/root/repo/piccolo_b200/csrc/pcl_sampling.o
for k in "${@:-ILi5ELb0E ILi5ELb1E}"; do for kk in $k; do
cuobjdump -sass -fun "_Z17pcl_sample_kernel${kk}Ev12PclCloudView8PclImagePKfiixPdPj11PclFinalize" $O | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/w/k_$kk.sass
python - "$kk" <<'PY'
import re,sys,collections
kk=sys.argv[1]
lines=open(f'/tmp/w/k_{kk}.sass').read().splitlines()
addr=[int(re.match(r"\s+/\*([0-9a-f]{4})\*/",l).group(1),16) for l in lines]
for i,l in enumerate(lines):
    m=re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)",l)
    if m:
        t=int(m.group(1),16)
        if t<addr[i]:
            n=(addr[i]-t)//16+1
            if 380<n<900:
                c=collections.Counter()
                for a,ll in zip(addr,lines):
                    if t<=a<=addr[i]:
                        mm=re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)",ll); c[mm.group(1)]+=1
                print(kk,"pose loop:",n,"instr (K=4: /4, K=5: /5 per eval)", sorted(c.items(), key=lambda x:-x[1])[:14])
PY
done; done
