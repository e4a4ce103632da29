"""Per-layer table from a bench.py --profile-out JSON (ResNet-50 naming)."""
import json, sys
def names_resnet50():
    order=[]; spec=[(64,64,256,3,1),(256,128,512,4,2),(512,256,1024,6,2),(1024,512,2048,3,2)]; hw=56
    for L,(i,m,o,n,s) in enumerate(spec,1):
        for bi in range(n):
            st=s if bi==0 else 1; ic=i if bi==0 else o
            if bi==0: order.append(f"L{L}.{bi}.ds 1x1/{st} {ic}->{o} @{hw}")
            order.append(f"L{L}.{bi}.c1 1x1 {ic}->{m} @{hw}")
            order.append(f"L{L}.{bi}.c2 3x3/{st} {m}->{m} @{hw}")
            hw2=hw//st
            order.append(f"L{L}.{bi}.c3 1x1 {m}->{o}+res @{hw2}")
            hw=hw2
    return order
def main(path, verbose=False):
    b=json.load(open(path))['launches']
    convs=[q for q in b if q['kind']=='conv_igemm']
    tot=sum(q['ms'] for q in b)
    agg={}
    for nm,q in zip(names_resnet50(),convs):
        key=nm.split('.')[0]+'.'+nm.split('.')[2].split(' ')[0]
        a=agg.setdefault(key,[0,0,0]); a[0]+=q['ms']; a[1]+=q['flops']; a[2]+=q['bytes']
        if verbose: print(f"{nm:34s} {q['ms']*1e3:7.1f} us {q['flops']/(q['ms']*1e-3)/1e12:7.1f} TF/s {q['bytes']/(q['ms']*1e-3)/1e9:6.0f} GB/s")
    for k,v in agg.items():
        ideal=max(v[1]/1385.5e12, v[2]/6544e9)*1e6
        print(f"{k:8s} {v[0]*1e3:7.1f} us  {100*v[0]/tot:4.1f}%  {v[1]/(v[0]*1e-3)/1e12:6.1f} TF/s {v[2]/(v[0]*1e-3)/1e9:6.0f} GB/s   roofline-ideal {ideal:6.1f} us  eff {ideal/(v[0]*1e3):.2f}")
    other=[q for q in b if q['kind']!='conv_igemm']
    for q in other: print(f"{q['kind']:10s} {q['ms']*1e3:7.1f} us  {100*q['ms']/tot:4.1f}%")
    print("total", tot*1e3, "us; conv", sum(q['ms'] for q in convs)*1e3)
if __name__=="__main__":
    main(sys.argv[1], len(sys.argv)>2)
