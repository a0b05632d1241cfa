"""Diagnostic: per-element comparison of D gradients / parameters after each of two training steps, GPU fp32
mode vs the CPU oracle on the tests' small config."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import sradsgan_oracle as O
from sradsgan_b200 import ops
from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup, SRADSGAN
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_model_parity import _args

golden = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/sradsgan_golden.pt"), weights_only=False)
gcfg = golden["train_steps"]["cfg"]
ng, nb, scale = gcfg["n_groups"], gcfg["n_blocks"], gcfg["scale"]
mk = lambda: (O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=gcfg["gseed"], init="fan")),
              O.make_state(O.discriminator_spec(), seed=gcfg["dseed"], init="ref"), O.make_state(O.vgg_spec(), seed=gcfg["vseed"], init="fan"))
Gsd, Dsd, Vsd = mk()
net = SRADSGAN(_args(vgg_state=Vsd, precision="fp32"))
net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=scale)
net.build(init=False)
net.generator.load_state_dict(Gsd, strict=True); net.discriminator.load_state_dict(Dsd, strict=True)
ops.bump_weight_generation()
G2, D2, V2 = mk()
st = O.TrainState(G2, D2, V2, scale, ng, nb)
for it in range(2):
    lr, hr = O.synthetic_batch(gcfg["batch"], scale, gcfg["lr_size"] * scale, seed=gcfg["data_seed"] + it)
    np.random.seed(gcfg["np_seed"] + it)
    alpha = torch.Tensor(np.random.random((gcfg["batch"], 1, 1, 1)))
    net._alpha_override = alpha
    out = net.train_step(lr.cuda(), hr.cuda())
    ref = O.train_step(st, lr, hr, alpha)
    print("step", it, {k: (out[k].item(), ref[k]) for k in ("loss_G", "loss_D", "gp")})
    dsd = net.discriminator.state_dict()
    for k, v in D2.items():
        if v.dtype.is_floating_point and v.dim() >= 1:
            a = dsd[k].float().cpu(); b = v.detach()
            ga = dict(net.discriminator.named_parameters()).get(k)
            gerr = ""
            if ga is not None and ga.grad is not None and b.grad is not None:
                gg = ga.grad.float().cpu(); rg = b.grad
                gerr = " grad rel %.2e (|g| %.2e)" % (((gg - rg).norm() / rg.norm().clamp_min(1e-30)).item(), rg.norm().item())
                if k == "model.3.bias":
                    print("   ours grad", gg[:8].tolist()); print("   ref  grad", rg[:8].tolist())
                    print("   ours p", a[:8].tolist()); print("   ref  p", b[:8].tolist())
            print("  %-28s rel %.2e norm %.3e%s" % (k, ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), b.norm().item(), gerr))
