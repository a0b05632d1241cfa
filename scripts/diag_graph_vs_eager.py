"""Diagnostic: eager train_step vs graphed_step on identical nets/inputs/alpha; prints every loss component per step."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sradsgan_oracle as O
from sradsgan_b200 import ops
from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup, SRADSGAN

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_model_parity import _args

ng, nb, scale, batch, lrs = 2, 1, 4, 2, 8
Gsd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=21, init="fan"))
Dsd = O.make_state(O.discriminator_spec(), seed=22, init="ref")
Vsd = O.make_state(O.vgg_spec(), seed=23, init="fan")
nets = []
for _ in range(3):
    net = SRADSGAN(_args(vgg_state=Vsd, precision="bf16"))
    net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=scale)
    net.build(init=False)
    net.generator.load_state_dict(Gsd, strict=True)
    net.discriminator.load_state_dict(Dsd, strict=True)
    ops.bump_weight_generation()
    nets.append(net)
batches = [O.synthetic_batch(batch, scale, lrs * scale, seed=31 + it) for it in range(3)]
keys = ("loss_G", "loss_D", "gp", "pixel", "content", "adv")
for name, net in zip(("eager", "eager2", "graph"), nets):
    for it, (lr, hr) in enumerate(batches):
        np.random.seed(77 + it)
        if name != "graph":
            net._alpha_override = torch.Tensor(np.random.random((batch, 1, 1, 1)))
            o = net.train_step(lr.cuda(), hr.cuda())
        else:
            o = net.graphed_step(lr.cuda(), hr.cuda())
        print(name, it, {k: round(o[k].item(), 6) for k in keys if k in o}, flush=True)
    print(name, "G", net.optimizer_G.flat_param.double().norm().item(), "D", net.optimizer_D.flat_param.double().norm().item())
e, g = nets[0], nets[2]
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
print("rel G", rel(g.optimizer_G.flat_param, e.optimizer_G.flat_param), "rel D", rel(g.optimizer_D.flat_param, e.optimizer_D.flat_param))
print("rel G eager2", rel(nets[1].optimizer_G.flat_param, e.optimizer_G.flat_param), "rel D eager2", rel(nets[1].optimizer_D.flat_param, e.optimizer_D.flat_param))
