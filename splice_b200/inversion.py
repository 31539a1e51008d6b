"""Drop-in for the reference's inversion.py (feature inversion: optimise a generator so that a DINO-ViT feature of its output
matches the feature of a given image). Same command line, same loop (ref inversion.py:12-74):

    net          = skip(32, 3, 6 scales, 7x7/5x5/3x3 filters, pad='reflection')   -> native engine csrc/generator_x.cu
    feature(x)   = vit_extractor.get_feature_from_input(pre(x))[layer][:, 0, :]    ('cls')
                 | vit_extractor.get_keys_from_input(pre(x), layer)               ('keys') -> native ViT engine, differentiable taps
    loss         = MSELoss(feature(net(net_input)), feature(image)); Adam(lr = 0.01) on net.parameters()

The torchvision transforms (Resize(224) + Normalize) stay torch ops, as in the reference: they are differentiable tensor ops
between the two native engines. Extensions (keyword arguments of `invert`, not on the command line): `vit_state_dict` - DINO weights
to use instead of torch.hub (offline runs), `callback(i, loss, net, net_input)` - called every iteration with the detached loss,
`prefetch_noise` (default on) - the 'cls' mode's per-iteration noise is drawn ahead by a worker thread (same random stream as the
reference's inline CPU draw, see NoiseFeed), `noise_on_device` - draw it with the CUDA generator instead (a different random stream
under the same seed; no host work at all).
"""
from __future__ import annotations

import queue
import threading
from argparse import ArgumentParser

import torch
from PIL import Image
from torchvision import transforms as T

from .models.extractor import VitExtractor
from .models.unet.skip import skip

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

# network configuration of the script (ref inversion.py:21-25); the input depth comes from the command line (default 32)
NET_ARGS = dict(num_channels_down=[16, 32, 64, 128, 128, 128],
                num_channels_up=[16, 32, 64, 128, 128, 128],
                num_channels_skip=[4, 4, 4, 4, 4, 4],
                filter_size_down=[7, 7, 5, 5, 3, 3], filter_size_up=[7, 7, 5, 5, 3, 3],
                downsample_mode='stride', pad='reflection')


class NoiseFeed:
    """The per-iteration regularisation noise of the 'cls' mode (ref inversion.py:56-62: `torch.randn(shape).to(device)` inside
    the loop - 2.1 M CPU normal draws, ~10 ms at 224 x 298, longer than the whole GPU iteration here), drawn `depth` iterations
    ahead by one worker thread into pinned memory. The stream is unchanged: the worker draws from torch's process-wide CPU generator
    in order, and nothing else draws from it inside the loop, so iteration i receives exactly the tensor the reference's inline
    call would have produced under the same seed. (What differs: when the loop ends the generator has advanced by up to `depth`
    extra draws.)"""

    def __init__(self, shape, n: int, depth: int = 4, pin=None):
        self.shape, self.n = tuple(shape), n
        self.pin = torch.cuda.is_available() if pin is None else pin
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._stop = threading.Event()
        self._worker = threading.Thread(target=self._run, name="splice-inversion-noise", daemon=True)
        self._worker.start()

    def _run(self):
        try:
            for _ in range(self.n):
                if self._stop.is_set():
                    return
                z = torch.randn(self.shape)
                if self.pin:
                    z = z.pin_memory()
                while not self._stop.is_set():
                    try:
                        self._q.put(z, timeout=0.1)
                        break
                    except queue.Full:
                        continue
        except BaseException as e:  # noqa: BLE001 - surfaced to the consumer
            self._q.put(e)

    def next(self) -> torch.Tensor:
        z = self._q.get()
        if isinstance(z, BaseException):
            raise z
        return z

    def close(self):
        self._stop.set()
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass
        self._worker.join(timeout=5)


def invert(args, vit_state_dict=None, callback=None, noise_on_device=False, prefetch_noise=True):
    if device.type != "cuda":
        raise RuntimeError("splice_b200 runs on sm_100a only; there is no CPU fallback")
    # load the image
    input_img = Image.open(args.image_path).convert("RGB")
    input_img = T.Compose([
        T.Resize(224),
        T.ToTensor()
    ])(input_img).unsqueeze(0).to(device)

    # network configurations (ref inversion.py:21-25)
    net = skip(args.input_depth, 3, **NET_ARGS).to(device)
    net_input_saved = torch.randn((1, args.input_depth, input_img.shape[-2], input_img.shape[-1])).to(device)

    # define the extractor
    dino_preprocess = T.Compose([
        T.Resize(224),
        T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    ])
    vit_extractor = VitExtractor(args.dino_model_name, device, state_dict=vit_state_dict)

    def extract_feature(x):
        if args.feature == 'cls':
            f = vit_extractor.get_feature_from_input(dino_preprocess(x))[args.layer][:, 0, :]
        elif args.feature == 'keys':
            f = vit_extractor.get_keys_from_input(dino_preprocess(x), args.layer)
        else:
            raise ValueError('feature {} not supported.'.format(args.feature))
        return f

    # calculate the target feature from the input image
    with torch.no_grad():
        ref_feature = extract_feature(input_img)

    # optimization configurations
    optimizer = torch.optim.Adam(net.parameters(), lr=args.LR)
    criterion = torch.nn.MSELoss()

    # inversion loop
    losses = []
    feed = None
    if args.feature == 'cls' and prefetch_noise and not noise_on_device:
        feed = NoiseFeed(net_input_saved.shape, args.n_iter)
    try:
        return _loop(args, net, net_input_saved, extract_feature, ref_feature, optimizer, criterion, losses, feed, noise_on_device,
                     callback)
    finally:
        if feed is not None:
            feed.close()


def _loop(args, net, net_input_saved, extract_feature, ref_feature, optimizer, criterion, losses, feed, noise_on_device, callback):
    for i in range(args.n_iter):
        net_input = net_input_saved
        if args.feature == 'cls':
            # noise added to the input at each step as a regularization (ref inversion.py:56-62)
            if noise_on_device:
                noise = torch.randn(net_input_saved.shape, device=device)
            elif feed is not None:
                noise = feed.next().to(device, non_blocking=True)
            else:
                noise = torch.randn(net_input_saved.shape).to(device)
            if i < args.reduce_noise_stage_1_iter:
                net_input = net_input_saved + (noise * 10)
            elif i < args.reduce_noise_stage_2_iter:
                net_input = net_input_saved + (noise * 2)
            else:
                net_input = net_input_saved + (noise * 0.5)

        optimizer.zero_grad()
        current_feature = extract_feature(net(net_input))

        loss = criterion(current_feature, ref_feature)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach())
        if callback is not None:
            callback(i, losses[-1], net, net_input)

        if i % args.log_freq == 0:
            with torch.no_grad():   # the reference runs this extra forward with autograd on and drops the graph
                result_img = net(net_input)[0].detach().cpu().clone()
            result_img = T.ToPILImage()(result_img)
            result_img.save(args.save_path)
    return net, torch.stack(losses).cpu() if losses else torch.empty(0)


# the reference script's command line (inversion.py:77-92): flag, type, default, meaning
_CLI = (
    ("feature", str, None, "which DINO-ViT feature of the image to invert: 'cls' ([CLS] token) or 'keys'"),
    ("layer", int, 11, "transformer block the feature is taken from (0-11)"),
    ("dino_model_name", str, "dino_vitb8", "dino_vits16 | dino_vits8 | dino_vitb16 | dino_vitb8"),
    ("image_path", str, "datasets/feature_visualization/limes.jpeg", "image whose feature is inverted"),
    ("save_path", str, None, "where the current result image is written every --log_freq iterations (required)"),
    ("log_freq", int, 100, "iterations between two result images"),
    ("input_depth", int, 32, "channels of the fixed noise tensor the generator maps to an image"),
    ("LR", float, 0.01, "Adam learning rate"),
    ("n_iter", int, 20000, "optimisation iterations"),
    ("reduce_noise_stage_1_iter", int, 10000, "'cls' mode: input noise of scale 10 before this iteration ..."),
    ("reduce_noise_stage_2_iter", int, 15000, "... of scale 2 before this one, 0.5 afterwards"),
)


def build_parser() -> ArgumentParser:
    parser = ArgumentParser(description="DINO-ViT feature inversion on the splice_b200 engines (same options as the reference script)")
    for flag, kind, default, text in _CLI:
        if flag == "save_path":
            parser.add_argument("--" + flag, type=kind, required=True, help=text)
        else:
            parser.add_argument("--" + flag, type=kind, default=default, help=text)
    return parser


if __name__ == '__main__':
    invert(build_parser().parse_args())
