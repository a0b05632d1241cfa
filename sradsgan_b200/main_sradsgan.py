"""CLI with the flags of the reference's main_sradsgan.py:16-63 (same names and defaults), plus
--precision / --synthetic_steps / --mode.  One process per GPU: launch with torchrun for data parallelism.

    python -m sradsgan_b200.main_sradsgan --scale_factor 4 --synthetic_steps 100 --num_epochs 1
"""
import argparse
import os

import torch


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="B200-native SRADSGAN")
    parser.add_argument('--model_name', type=str, default='SRADSGAN', choices=['SRADSGAN', 'EDSR'], help='The type of model')
    parser.add_argument('--root_dir', type=str, default='./')
    parser.add_argument('--data_dir', type=str, default='./dataset/sradsgan/')
    parser.add_argument('--train_dataset', type=list, default=["AID", "DOTA", "LoveDA", "RSSCN7_2800", "SECOND"])
    parser.add_argument('--test_dataset', type=list, default=["UCMerced_LandUse"])
    parser.add_argument('--crop_size', type=int, default=216, help='Size of cropped HR image')
    parser.add_argument('--num_threads', type=int, default=16, help='number of threads for data loader to use')
    parser.add_argument('--num_channels', type=int, default=3, help='The number of channels to super-resolve')
    parser.add_argument('--scale_factor', type=int, default=8, help='Size of scale factor')
    parser.add_argument('--epoch', type=int, default=0, help='epoch to start training from')
    parser.add_argument('--num_epochs', type=int, default=100, help='The number of epochs to run')
    parser.add_argument('--save_epochs', type=int, default=1, help='Save trained model every this epochs')
    parser.add_argument('--batch_size', type=int, default=16, help='training batch size (per GPU)')
    parser.add_argument('--test_batch_size', type=int, default=1, help='testing batch size')
    parser.add_argument('--save_dir', type=str, default='Result', help='Directory name to save the results')
    parser.add_argument('--lr', type=float, default=0.0002, help='learning rate default 0.0002')
    parser.add_argument('--b1', type=float, default=0.9, help='adam: decay of first order momentum of gradient')
    parser.add_argument('--b2', type=float, default=0.999, help='adam: decay of first order momentum of gradient')
    parser.add_argument('--gpu_mode', type=bool, default=True)
    parser.add_argument('--test_crop_size', type=int, default=216, help='Size of cropped HR image')
    parser.add_argument('--n_cpu', type=int, default=16, help='number of cpu threads to use during batch generation')
    parser.add_argument('--hr_height', type=int, default=216, help='size of high res. image height')
    parser.add_argument('--hr_width', type=int, default=216, help='size of high res. image width')
    parser.add_argument('--sample_interval', type=int, default=1000, help='interval between sampling of images from generators')
    parser.add_argument('--clip_value', type=float, default=0.01, help='lower and upper clip value for disc. weights')
    parser.add_argument('--lambda_gp', type=float, default=10, help='Loss weight for gradient penalty')
    parser.add_argument('--gp', type=bool, default=True, help='gradient penalty')
    parser.add_argument('--penalty_type', type=str, default='LS', choices=["LS", "hinge"], help='gradient type')
    parser.add_argument('--grad_penalty_Lp_norm', type=str, default='L2', choices=["L2", "L1", "Linf"], help='gradient penalty Lp norm')
    parser.add_argument('--relativeGan', type=bool, default=False, help='relative GAN')
    parser.add_argument('--loss_Lp_norm', type=str, default='L1', choices=["L2", "L1"], help='loss Lp norm')
    parser.add_argument('--weight_content', type=float, default=1e-2, help='Loss weight for content loss')
    parser.add_argument('--weight_gan', type=float, default=1e-3, help='Loss weight for gan loss')
    parser.add_argument('--max_train_samples', type=int, default=40000, help='Max training samples')
    parser.add_argument('--is_train', type=bool, default=True, help='if at training stage')
    # new
    parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32'])
    parser.add_argument('--synthetic_steps', type=int, default=0, help='>0: train on synthetic batches, this many per epoch')
    parser.add_argument('--log_interval', type=int, default=50)
    parser.add_argument('--vgg_state', type=str, default=None, help='state_dict file for VGG19 features[:12]')
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--mode', type=str, default='train', choices=['train', 'validate', 'test_single'])
    parser.add_argument('--img', type=str, default=None)
    parser.add_argument('--modelpath', type=str, default=None)
    parser.add_argument('--tile', type=int, default=0)
    parser.add_argument('--gpu_input_pipeline', action='store_true', help='decode+crop on the workers, PIL-exact bicubic LR synthesis and copies on the device')
    parser.add_argument('--pretrained_G', type=str, default=None, help='generator state_dict to warm-start from (chain training)')
    parser.add_argument('--pretrained_D', type=str, default=None, help='discriminator state_dict to warm-start from')
    parser.add_argument('--no_graphs', dest='graphs', action='store_false', help='launch every kernel of an iteration from Python instead of replaying the captured CUDA graph (default: graphs on)')
    parser.add_argument('--chain_scales', type=str, default='', help="e.g. '2,3,4': train these scales one after the other, each warm-started from the previous")
    return check_args(parser.parse_args(argv))


def check_args(args):
    """reference main_sradsgan.py:66-86"""
    from .utils import mkdir_and_rename
    args.save_dir = os.path.join(args.save_dir, args.model_name)
    if args.epoch == 0 and args.mode == 'train' and int(os.environ.get("RANK", "0")) == 0:
        mkdir_and_rename(os.path.join(args.root_dir, args.save_dir))
    if args.num_epochs < 1:
        print('number of epochs must be larger than or equal to one')
    if args.batch_size < 1:
        print('batch size must be larger than or equal to one')
    return args


def main(argv=None):
    args = parse_args(argv)
    if args.gpu_mode and not torch.cuda.is_available():
        raise Exception("No GPU found, please run without --gpu_mode=False")
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    from .model.sradsgan import SRADSGAN
    if args.model_name == 'EDSR':
        from .model.edsr import EDSR
        net = EDSR(args)
    elif args.model_name == 'SRAGAN':
        from .model.sragan import SRAGAN        # reference main_sragan.py: SRADSGAN's predecessor
        net = SRAGAN(args)
    elif args.model_name == 'NDSRGAN':
        from .model.ndsrgan import NDSRGAN      # reference main_ndsrgan.py
        net = NDSRGAN(args)
    elif args.model_name == 'SRGAN':
        from .model.srgan import SRGAN          # reference main_srgan.py: the first of the sibling GANs
        net = SRGAN(args)
    else:
        net = SRADSGAN(args)
    if args.mode == 'train' and args.chain_scales:
        net.chain_train([int(v) for v in args.chain_scales.split(',')])
    elif args.mode == 'train':
        net.train()
    elif args.mode == 'validate':
        print(net.mfeNew_validateByClass(100, save_img=True, modelpath=args.modelpath))
    else:
        net.mfe_test_single(img_fn=args.img, modelpath=args.modelpath, tile=args.tile or None)


if __name__ == '__main__':
    main()
