import sys, torch, time
sys.path.insert(0,'vox-e_b200')
from voxe_b200 import _native as nat
from voxe_b200.render_function import FusedGridSpec, pack_volume
lib = nat.load_library()
dev = torch.device('cuda')
dims=(160,160,160)
dens = torch.rand(*dims,1,device=dev); feat = torch.rand(*dims,3,device=dev)
spec = FusedGridSpec(dims=dims,n_features=3,aabb=((-1.5,1.5),)*3,density_scale=1.0,preact=0,postact=1)
gd = spec.to_native()
packed = pack_volume(spec,dens,feat); pg = torch.randn_like(packed); m = torch.zeros_like(packed); v = torch.zeros_like(packed)
adam = nat.VoxeAdamDesc(lr=0.03,beta1=0.9,beta2=0.999,eps=1e-8,step=1)
s = torch.cuda.current_stream().cuda_stream
def run(n):
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        lib.voxe_adam_step(gd, adam, dens.data_ptr(), feat.data_ptr(), packed.data_ptr(), pg.data_ptr(), None, None, m.data_ptr(), v.data_ptr(), s)
    b.record(); torch.cuda.synchronize(); return 1e3*a.elapsed_time(b)/n
run(3); print('adam kernel us', run(20))
t0=time.perf_counter(); 
for _ in range(100): lib.voxe_adam_step(gd, adam, dens.data_ptr(), feat.data_ptr(), packed.data_ptr(), pg.data_ptr(), None, None, m.data_ptr(), v.data_ptr(), s)
print('cpu us/call', (time.perf_counter()-t0)/100*1e6); torch.cuda.synchronize()
# copy bandwidth reference
x = torch.empty(2**28, device=dev); y = torch.empty_like(x)
def cp(n):
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): y.copy_(x)
    b.record(); torch.cuda.synchronize(); return 1e3*a.elapsed_time(b)/n
cp(3); t=cp(10); print('copy GB/s', 2*x.numel()*4/t/1e3)

# --- representative measurement: refill the gradient volume (as a backward pass would) before every step, in-stream
noise = torch.randn_like(packed)
def refill_and_step(n, step=True):
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        pg.copy_(noise)
        if step:
            lib.voxe_adam_step(gd, adam, dens.data_ptr(), feat.data_ptr(), packed.data_ptr(), pg.data_ptr(), None, None, m.data_ptr(), v.data_ptr(), s)
    b.record(); torch.cuda.synchronize(); return 1e3*a.elapsed_time(b)/n
refill_and_step(3); both = refill_and_step(20); only = refill_and_step(20, step=False)
print('refill+step us', both, 'refill us', only, '=> adam kernel on real data us', both-only)
sys.path.insert(0,'.')
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
from voxe_b200.optim import FusedVoxelAdam
grid = VoxelGrid(dens.clone(), feat.clone(), VoxelSize(3/160,3/160,3/160), density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(), tunable=True)
grid.packed_cache().get(grid.fused_spec(), grid.densities, grid.features)
opt = FusedVoxelAdam(grid, lr=0.03); acc = grid.render_gradient_accumulator; buf = acc.get(packed)
for _ in range(3): buf.copy_(noise); acc.dirty=True; opt.step()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(50): acc.dirty=True; opt.step()
cpu = (time.perf_counter()-t0)/50*1e6; torch.cuda.synchronize(); print('FusedVoxelAdam.step CPU us/call (async)', cpu)
