"""Hot source lines of an .ncu-rep (read here, no GPU): % of stall samples and of warp instructions per CUDA source line.
    python tools/ncu_toplines.py gpurun_out/x.ncu-rep [N]"""
import csv,sys,subprocess,collections
rep=sys.argv[1]; N=int(sys.argv[2]) if len(sys.argv)>2 else 30
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass,cuda"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
fp=None; res=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": fp=r[1].split('/')[-1]; continue
    if r[0] in ("Function Name","Line No"): continue
    if r[0]!="":
        try: res.append((int(r[6]),int(r[7]),fp,r[0],r[1].strip()[:105]))
        except ValueError: pass
res.sort(reverse=True)
ts=sum(o[0] for o in res); ti=sum(o[1] for o in res)
print(ts,ti)
for o in res[:N]: print("%5.1f%% %5.1f%%"%(100*o[0]/ts,100*o[1]/ti),o[2:])
