import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, tneq_b200
from oracle import qctn_oracle as oc
from helpers import make_case, clone_mx
H = tneq_b200.QCTNHelper
def graph_of(kind,n,K):
    if kind=='merged':
        q = tneq_b200.QCTN(H.generate_example_graph(n=n, graph_type='mps', dim_char=str(K))); return tneq_b200.QCTN.merge(q,q).graph
    return H.generate_example_graph(n=n, graph_type=kind, dim_char=str(K))
def up(x, td):
    if isinstance(x, oc.TNT): return oc.TNT(x.tensor.to(td), x.scale, x.log_scale)
    return x.to(td)
def todev(x):
    if isinstance(x, oc.TNT): return tneq_b200.TNTensor(x.tensor.cuda(), x.scale, x.log_scale)
    return x.cuda()
def mx(a,b): return ((a-b).abs().max()/b.abs().max()).item()
for kind,n,K,B,dtype in [('mps',6,3,64,'float32'),('mps',16,3,300,'float32'),('mps',16,3,4096,'float32'),('tree',7,3,21,'float32'),('merged',6,3,40,'float32'),('mps',6,3,19,'complex64'),('merged',4,2,9,'complex64'),('mps',24,3,512,'float32')]:
    graph = graph_of(kind,n,K)
    names, table, nq, cores, states, mxs = make_case(graph, K, B, dtype, tnt=True)
    td64 = torch.float64 if dtype=='float32' else torch.complex128
    ref = oc.forward(graph, cores, states, clone_mx(mxs))
    lref, gref = oc.loss_and_grads(graph, cores, states, clone_mx(mxs))
    c64 = {k: v.to(td64) for k,v in cores.items()}; s64=[s.to(td64) for s in states]
    tru = oc.forward(graph, c64, s64, [up(m,td64) for m in clone_mx(mxs)])
    ltru, gtru = oc.loss_and_grads(graph, c64, s64, [up(m,td64) for m in clone_mx(mxs)])
    be = tneq_b200.BackendFactory.create_backend('b200', device='cuda:0', dtype=dtype)
    eng = tneq_b200.EngineSiamese(backend=be, strategy_mode='balanced', mx_K=K)
    q = tneq_b200.QCTN(graph, backend=be)
    for k,v in cores.items(): q.cores_weights[k] = v.cuda().requires_grad_(True)
    st = [s.cuda() for s in states]
    got = eng.contract_with_compiled_strategy(q, st, [todev(m) for m in clone_mx(mxs)]).cpu()
    loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, st, [todev(m) for m in clone_mx(mxs)])
    grads=[g.cpu() for g in grads]
    pe = lambda a,b: (((a-b).abs()/b.abs()).max().item())
    print(f"{kind}{n} K{K} B{B} {dtype}: P maxnorm ours-ref {mx(got,ref):.1e} ours-tru {mx(got.double(),tru):.1e} ref-tru {mx(ref.double(),tru):.1e} | P elem ours-tru {pe(got.double(),tru):.1e} ref-tru {pe(ref.double(),tru):.1e} | loss ours-ref {abs(loss.item()-lref.item())/abs(lref.item()):.1e} ours-tru {abs(loss.item()-ltru.item())/abs(ltru.item()):.1e} ref-tru {abs(lref.item()-ltru.item())/abs(ltru.item()):.1e} | G ours-ref {max(mx(a,b) for a,b in zip(grads,gref)):.1e} ours-tru {max(mx(a.to(td64),b) for a,b in zip(grads,gtru)):.1e} ref-tru {max(mx(a.to(td64),b) for a,b in zip(gref,gtru)):.1e}")
