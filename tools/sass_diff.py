"""Per-function SASS comparison of two builds of libdmb200.so (is the hot kernel untouched by a host-side or
add-only change?):  python tools/sass_diff.py old.so new.so"""
import subprocess, sys, hashlib, re
def funcs(path):
    out = subprocess.run(["cuobjdump","-sass",path],capture_output=True,text=True).stdout
    res={}; cur=None
    for line in out.splitlines():
        m=re.match(r"\s*Function : (\S+)", line)
        if m: cur=m.group(1); res[cur]=[]; continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            res[cur].append(re.sub(r"/\*[0-9a-f]{16}\*/","",line).strip())
    return {k:hashlib.md5("\n".join(v).encode()).hexdigest() for k,v in res.items()}
a=funcs(sys.argv[1]); b=funcs(sys.argv[2])
same=[k for k in a if k in b and a[k]==b[k]]; diff=[k for k in a if k in b and a[k]!=b[k]]
print("functions old/new:",len(a),len(b),"identical:",len(same),"changed:",len(diff))
for k in diff: print("CHANGED",k)
print("new only:",[k for k in b if k not in a])
