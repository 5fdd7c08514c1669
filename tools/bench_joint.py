"""BASELINE config 4: stage-2 joint SR + segmentation training step with the UASR teacher and structural knowledge
distillation, batch 2 per GPU, data parallel (torchrun for N > 1).  SURVEY.md 8(d) C4 protocol: anisotropic SegModel student
(upscale 4), img ~ N(0,1) [2,1,16,256,256], label_lr / label_sr = (U > 0.8), uncertainty ~ U(0.01, 1), SGD(momentum 0.99,
nesterov, wd 3e-5) as train_all.py:513.  Prints one JSON line (device time, max over ranks)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from rehrseg_b200 import flavr, functional as Fn, loss_ops, seg_model as sm, train_step as ts

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--no-distill", action="store_true")
ap.add_argument("--breakdown", action="store_true")
ap.add_argument("--plain-losses", action="store_true", help="plain-PyTorch Distiller / losses instead of the fused kernels")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

ANISO = dict(input_channels=1, n_stages=6, features_per_stage=[32, 64, 128, 256, 320, 320], conv_op=torch.nn.Conv3d,
             kernel_sizes=[[1, 3, 3], [1, 3, 3]] + [[3, 3, 3]] * 4,
             strides=[[1, 1, 1], [1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2], [1, 2, 2]], n_conv_per_stage=[2] * 6, num_classes=2,
             upscale=4, n_conv_per_stage_decoder=[2] * 5, conv_bias=True, norm_op=torch.nn.InstanceNorm3d,
             norm_op_kwargs={"eps": 1e-5, "affine": True}, dropout_op=None, dropout_op_kwargs=None, nonlin=torch.nn.LeakyReLU,
             nonlin_kwargs={"inplace": True}, deep_supervision=False)
torch.manual_seed(1234)
student = sm.SegModel(**ANISO).to(dev)
teacher = flavr.UNet_3D_3D(2, "unet_18", 4, 4, False, "concat", "transpose", use_uncertainty=True).to(dev).eval()
distiller = (ts.Distiller if a.plain_losses else loss_ops.FusedDistiller)(64, 64, 0.0, 1.0, 1.0).to(dev)
import itertools
opt = torch.optim.SGD(itertools.chain(student.parameters(), distiller.parameters()), lr=1e-3, momentum=0.99, nesterov=True,
                      weight_decay=3e-5)
build = ts.build_loss if a.plain_losses else loss_ops.build_fused_loss
lr_obj, hr_obj = build(False, 0), build(False, 1)

g = torch.Generator().manual_seed(4 + rank)
B, D, HW = 2, 16, 256
host = (torch.randn((B, 1, D, HW, HW), generator=g).pin_memory(),
        (torch.rand((B, 1, D, HW, HW), generator=g) > 0.8).float().pin_memory(),
        (torch.rand((B, 1, 4 * D, HW, HW), generator=g) > 0.8).float().pin_memory(),
        (torch.rand((B, 1, D, HW, HW), generator=g) * 0.99 + 0.01).pin_memory())


def step():
    # host -> device copies of the batch are part of the step, as in train_all.py:524-541
    return ts.joint_train_step(student, host, lr_obj, hr_obj, opt, None if a.no_distill else teacher,
                               None if a.no_distill else distiller, device=dev)


for _ in range(a.warmup):
    out = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = Fn.launches()
e0.record()
for _ in range(a.steps):
    out = step()
    float(out["loss"])                     # the loop reads its loss back like a training script printing it
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
t = torch.tensor([ms], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t)
res = {"config": "C4 joint SR+seg step, student anisotropic SegModel [2,1,16,256,256] x4 SR head, UASR FLAVR teacher sweep (15 windows), "
                 "Distiller(64,64,0,1,1), SGD; batch 2 per GPU", "n_gpus": world, "ms_per_step": round(ms, 3),
       "samples_per_s": round(world * B / ms * 1e3, 2), "steps": a.steps, "warmup": a.warmup, "distill": not a.no_distill,
       "losses": "plain PyTorch (train_step.py)" if a.plain_losses else "fused CUDA reductions (loss_ops.py)",
       "engine_launches_per_step": (Fn.launches() - l0) // a.steps, "loss": {k: round(float(v), 5) for k, v in out.items()},
       "algorithmic_tflop_per_gpu_step": None if a.no_distill else 15.5}
if not a.no_distill:
    res["tflops_per_gpu"] = round(15.5 / ms * 1e3, 1)

if a.breakdown and rank == 0:
    def timed(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    dbatch = tuple(t.to(dev) for t in host)

    def teach():
        with torch.no_grad():
            return flavr.get_intermediate_features(teacher, dbatch[0].clone(), dbatch[1], dev, normalize=flavr.zscore_normalization)
    feats = teach()

    def stud_fwd():
        return student(dbatch[0], return_inetermediate_feature=True)

    def stud_fb():
        for p in student.parameters():
            p.grad = None
        o, u, s = stud_fwd()
        (o.float().mean() + u.float().mean() + s[1].float().mean()).backward()

    def losses():
        o, u, s = outs
        l = lr_obj(o, dbatch[1], dbatch[3]) + hr_obj(u, dbatch[2], None) + distiller(s[1], feats[1])
        return l
    with torch.no_grad():
        o, u, s = stud_fwd()
        outs = (o.detach().float().requires_grad_(True), u.detach().float().requires_grad_(True), [None, s[1].detach().float().requires_grad_(True)])

    def loss_fb():
        for t_ in (outs[0], outs[1], outs[2][1]):
            t_.grad = None
        losses().backward()
    res["breakdown_ms"] = {"teacher_sweep": round(timed(teach), 3), "student_fwd_bwd_mean_loss": round(timed(stud_fb), 3),
                           "losses_fwd_bwd": round(timed(loss_fb), 3), "h2d_batch": round(timed(lambda: [t.to(dev) for t in host]), 3)}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
