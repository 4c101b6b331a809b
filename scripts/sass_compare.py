import subprocess, re, sys, hashlib
def kernels(path):
    out = subprocess.run(["cuobjdump","-sass",path],capture_output=True,text=True).stdout
    res = {}; name=None; buf=[]
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name: res[name]="\n".join(buf)
            name=m.group(1); buf=[]
        elif name:
            # strip addresses/encodings comments
            l = re.sub(r"/\*[0-9a-fx]+\*/","",line).strip()
            if l: buf.append(l)
    if name: res[name]="\n".join(buf)
    return res
a=kernels(sys.argv[1]); b=kernels(sys.argv[2])
for k in sorted(set(a)|set(b)):
    d = subprocess.run(["cu++filt",k],capture_output=True,text=True).stdout.strip()[:110]
    if k not in a: print("NEW   ",d)
    elif k not in b: print("GONE  ",d)
    elif a[k]!=b[k]: print("DIFF  ",d)
    else: print("same  ",d)
