"""Fuzz the query server (loom_b200/host/loom_query.cc + the Query codec of loom_b200/io/lb_io.cc) with requests that
are structurally valid protobuf but carry hostile values: wrong-size masks, unsorted / out-of-range sparse lists,
out-of-range categories, non-finite reals, tare ids the store lacks, diffs that do not fit the tare, zero counts and
limits.  Every run must exit 0 with one response per request (errors are reported in Response.error).

Build the server against the oracle with sanitizers first (CPU only; the device library is not involved):

  cd /tmp/asan && R=/root/repo
  for f in loom_oracle oracle_math oracle_structure; do gcc -O1 -g -fsanitize=address,undefined -fopenmp -std=gnu11 -c $R/oracle/$f.c; done
  g++ -O1 -g -fsanitize=address,undefined -fopenmp -std=c++17 -c $R/loom_b200/io/lb_io.cc
  g++ -O1 -g -fsanitize=address,undefined -fopenmp -std=c++17 '-DLB_NAME(x)=lo_##x' $R/loom_b200/host/loom_query.cc \
      lb_io.o loom_oracle.o oracle_math.o oracle_structure.o -o loom_query_asan -lz -lpthread -lm

usage: fuzz_query.py SEED ITERATIONS [EXE] [tared]     (round 1: 4 seeds x 80 iterations on plain and tared stores, no
memory errors; two clean-abort classes found and fixed, see tests/test_query_server.py::test_hostile_*)"""
import sys, os, subprocess, pathlib, tempfile, struct, random
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_query_server as qs
from test_pbio import M
import pb_schema
R='/root/repo'
tmp=pathlib.Path(tempfile.mkdtemp())
subprocess.check_call(["/usr/bin/g++","-O1","-std=c++17","-DLB_NAME(x)=lo_##x",R+"/loom_b200/host/loom_infer.cc","-o",str(tmp/"li"),"-L"+R+"/oracle","-lloom_oracle","-L"+R+"/loom_b200/io","-lloomb200_io","-Wl,-rpath,"+R+"/oracle","-Wl,-rpath,"+R+"/loom_b200/io"])
root = tmp / "store"
model, block, rowids = qs.make_store(root, str(tmp/"li"), tared=len(sys.argv) > 4, n_rows=60)
F=model.n_features
rng=random.Random(int(sys.argv[1]))
def rand_observed(o):
    if rng.random()<0.7:
        feats=sorted(rng.sample(range(F), rng.randint(0,4)))
        if rng.random()<0.5: o.sparsity=1; o.sparse.extend(feats)
        else: o.sparsity=2; o.dense.extend([f in feats for f in range(F)])
        return
    o.sparsity=rng.choice([0,1,2,3])
    if rng.random()<0.7: o.dense.extend([rng.random()<0.5 for _ in range(rng.choice([F,F,F,F-1,F+1,0]))])
    if rng.random()<0.7: o.sparse.extend(sorted(rng.sample(range(F+2), rng.randint(0,5))) if rng.random()<0.8 else [5,3,9999])
from test_pbio import fill_value
def valid_value(v):
    cells={}
    for f in rng.sample(range(F), rng.randint(0,F)):
        t=model.features[f]["type"]
        if t=="bb": cells[f]=rng.random()<0.5
        elif t=="dd": cells[f]=rng.randrange(len(model.features[f]["alphas"])) if rng.random()<0.97 else 10**6
        elif t=="gp": cells[f]=rng.choice([0,1,2,5,40,10**5])
        elif t=="nich": cells[f]=rng.choice([0.0,1.5,-3.0,1e30,float('nan'),float('inf'),-1e-30])
        else: cells[f]=0
    fill_value(v, model, cells, rng.choice(["auto","dense","sparse"]) if cells and len(cells)<F else "auto")
def rand_value(v):
    if rng.random()<0.8: return valid_value(v)
    rand_observed(v.observed)
    v.booleans.extend([rng.random()<0.5 for _ in range(rng.randint(0,6))])
    v.counts.extend([rng.choice([0,1,3,15,255,256,10**6]) for _ in range(rng.randint(0,8))])
    v.reals.extend([rng.choice([0.0,1.5,-3.0,1e30,float('nan'),float('inf')]) for _ in range(rng.randint(0,4))])
def rand_diff(d):
    rand_value(d.pos)
    if rng.random()<0.85: d.neg.observed.sparsity=0
    else: rand_value(d.neg)
    if rng.random()<0.3: d.tares.extend([rng.choice([0,0,1,7])])
crashes=0; clean=0; ok=0; msgs={}
exe = sys.argv[3] if len(sys.argv) > 3 else '/tmp/asan/loom_query_asan'
env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
for it in range(int(sys.argv[2])):
    reqs=[]
    for i in range(4):
        r=qs.new_request(i)
        if rng.random()<0.6:
            rand_diff(r.score.data)
        if rng.random()<0.5:
            rand_diff(r.sample.data); rand_observed(r.sample.to_sample); r.sample.sample_count=rng.choice([0,1,3,50])
        if rng.random()<0.4:
            for _ in range(rng.randint(0,3)): rand_observed(r.entropy.row_sets.add())
            for _ in range(rng.randint(0,3)): rand_observed(r.entropy.col_sets.add())
            rand_diff(r.entropy.conditional); r.entropy.sample_count=rng.choice([0,1,2,20])
        if rng.random()<0.4:
            rand_diff(r.score_derivative.update_data)
            for _ in range(rng.randint(0,3)): rand_diff(r.score_derivative.score_data.add())
            r.score_derivative.row_limit=rng.choice([0,1,1000])
        reqs.append(r)
    pb_schema.dump_stream(tmp/"req.pbs", reqs)
    p=subprocess.run([exe,str(root),str(tmp/"req.pbs"),str(root/"query_config.pb.gz"),str(tmp/"resp.pbs"),"--none"],capture_output=True,env=env,timeout=300)
    err=p.stderr.decode(errors='replace')
    if "AddressSanitizer" in err or "runtime error" in err or p.returncode==-11:
        crashes+=1; print("CRASH", it, p.returncode, err[-2500:]); (tmp/"crash.pbs").write_bytes((tmp/"req.pbs").read_bytes()); break
    elif p.returncode!=0:
        clean+=1; k=err.split("\n")[0][:150]; msgs[k]=msgs.get(k,0)+1
    else:
        ok+=1
        resp=pb_schema.load_stream(tmp/"resp.pbs", M()["Query.Response"])
        assert len(resp)==4
        for x in resp:
            for e in x.error: msgs["resp: "+e]=msgs.get("resp: "+e,0)+1
print(msgs); print("ok",ok,"aborted",clean,"crashes",crashes,tmp)
