"""Config C3: one main-phase batch step of iMDBN.train_joint (represent + free CD-1 + aux clamped CD with
30 cond steps + _cross_reconstruct with 50 steps + metrics), batch 64."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_idbn_b200 as M
M.set_precision(os.environ.get("PREC", "tf32"))
dev = torch.device("cuda")
os.chdir("/tmp")
P = dict(LEARNING_RATE=0.1, WEIGHT_PENALTY=1e-4, INIT_MOMENTUM=0.5, FINAL_MOMENTUM=0.95, LEARNING_RATE_DYNAMIC=True,
         CD=1, JOINT_LEARNING_RATE=0.04, JOINT_CD=1, CROSS_GIBBS_STEPS=50, JOINT_AUX_COND_STEPS=30)
N, B = 64 * 8, 64
g = torch.Generator().manual_seed(0)
x = (torch.rand(N, 10000, generator=g) < 0.1).float()
y = torch.nn.functional.one_hot(torch.randint(0, 32, (N,), generator=g), 32).float()
dl = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x.pin_memory(), y.pin_memory()), batch_size=B)
m = M.iMDBN([10000, 1500, 500], 256, params=P, dataloader=dl, val_loader=None, device=dev, num_labels=32)
m.WARMUP_Y_EPOCHS = 0            # time the main phase
m.train_joint(1)                 # warm-up epoch (8 batches)
torch.cuda.synchronize()
t0 = time.perf_counter()
m.train_joint(3)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
nb = 3 * (N // B)
print(f"train_joint main phase: {dt/nb*1e3:.3f} ms per batch of {B}  ({nb*B/dt:.0f} samples/s)  metrics={m.metrics_history[-1]}")
