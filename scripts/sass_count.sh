#!/bin/bash
# Instruction count of the K=4 pose loop (per evaluation) of the u8q scoring and fwd+bwd kernels.
O=/root/repo/piccolo_b200/csrc/pcl_sampling.o
for k in "${@:-ILi1ELb0ELi4ELi0 ILi1ELb1ELi4ELi0}"; do for kk in $k; do
cuobjdump -sass -fun "_Z17pcl_sample_kernel${kk}EEv12PclCloudView8PclImagePKfiixPdPj11PclFinalize" $O | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/w/k_$kk.sass
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
                print(kk,"K=4 pose loop:",n,"->",n/4,"per eval", sorted(c.items(), key=lambda x:-x[1])[:14])
PY
done; done
