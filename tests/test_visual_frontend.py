"""Visual front-end: the batched trunk calls equal the reference's per-image / per-ROI loops
(run_multimodal_fcmf.py:449-460 over resnet_utils.py:6-55) -- restated below line by line -- on a small ResNet-shaped trunk."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from _util import pkg, rel_err

VF = pkg("visual_frontend")


class TinyResNet(nn.Module):
    """torchvision-ResNet attribute names, tiny widths."""

    def __init__(self, c=16):
        super().__init__()
        self.conv1 = nn.Conv2d(3, c, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(c)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=False), nn.BatchNorm2d(c), nn.ReLU())
        self.layer2 = nn.Sequential(nn.Conv2d(c, 2 * c, 3, 2, 1, bias=False), nn.BatchNorm2d(2 * c), nn.ReLU())
        self.layer3 = nn.Sequential(nn.Conv2d(2 * c, 4 * c, 3, 2, 1, bias=False), nn.BatchNorm2d(4 * c), nn.ReLU())
        self.layer4 = nn.Sequential(nn.Conv2d(4 * c, 8 * c, 3, 1, 1, bias=False), nn.BatchNorm2d(8 * c), nn.ReLU())


def ref_img(resnet, x, att_size=7):                         # myResNetImg.forward, resnet_utils.py:13-30
    x = resnet.maxpool(resnet.relu(resnet.bn1(resnet.conv1(x))))
    x = resnet.layer4(resnet.layer3(resnet.layer2(resnet.layer1(x))))
    return F.adaptive_avg_pool2d(x, [att_size, att_size]).detach()


def ref_roi(resnet, x):                                     # myResNetRoI.forward, resnet_utils.py:39-55
    x = resnet.maxpool(resnet.relu(resnet.bn1(resnet.conv1(x))))
    x = resnet.layer4(resnet.layer3(resnet.layer2(resnet.layer1(x))))
    return x.mean(3).mean(2).detach()


def reference_loop(img_net, roi_net, t_img, roi_img, C):    # run_multimodal_fcmf.py:445-460
    roi_img = roi_img.float()
    NI, NR = t_img.shape[1], roi_img.shape[2]
    encoded_img = [ref_img(img_net, t_img[:, i, :]).view(-1, C, 49).permute(0, 2, 1).squeeze(1) for i in range(NI)]
    encoded_roi = [torch.stack([ref_roi(roi_net, roi_img[:, i, r, :]).squeeze(1) for r in range(NR)], dim=1) for i in range(NI)]
    return torch.stack(encoded_img, dim=1), torch.stack(encoded_roi, dim=1)


def _data(B=3, NI=2, NR=2, hw=112):
    g = torch.Generator().manual_seed(0)
    return torch.randn(B, NI, 3, hw, hw, generator=g), torch.randn(B, NI, NR, 3, hw, hw, generator=g, dtype=torch.float64)


def test_batched_front_end_equals_reference_loops_eval_and_train():
    torch.manual_seed(1)
    img_net, roi_net = TinyResNet(), TinyResNet()
    t_img, roi_img = _data()
    for mode in ("eval", "train"):                          # train(): BatchNorm batch statistics -> the reference's call grouping is kept
        for net in (img_net, roi_net):
            net.train(mode == "train")
        sd_i = {k: v.clone() for k, v in img_net.state_dict().items()}
        sd_r = {k: v.clone() for k, v in roi_net.state_dict().items()}
        want_v, want_r = reference_loop(img_net, roi_net, t_img, roi_img, 128)
        img_net.load_state_dict(sd_i); roi_net.load_state_dict(sd_r)        # undo the running-stat updates of the reference pass
        fe = VF.VisualFrontEnd(img_net, roi_net, if_fine_tune=False)
        got_v, got_r = fe(t_img, roi_img)
        assert got_v.shape == want_v.shape == (3, 2, 49, 128) and got_r.shape == want_r.shape == (3, 2, 2, 128)
        assert rel_err(got_v, want_v) < 1e-5 and rel_err(got_r, want_r) < 1e-5, mode
        assert not got_v.requires_grad and not got_r.requires_grad
    assert sorted(k for k in fe.state_dict() if k.startswith("resnet_img.resnet.conv1"))     # reference checkpoint key layout


def test_feature_cache_and_fine_tune_gradients():
    torch.manual_seed(2)
    img_net, roi_net = TinyResNet().eval(), TinyResNet().eval()
    t_img, roi_img = _data()
    fe = VF.VisualFrontEnd(img_net, roi_net, cache=True)
    v1, r1 = fe(t_img, roi_img, keys=["a", "b", "c"])
    calls = {"n": 0}
    orig = img_net.conv1.forward
    img_net.conv1.forward = lambda x: (calls.__setitem__("n", calls["n"] + 1), orig(x))[1]
    v2, r2 = fe(t_img, roi_img, keys=["a", "b", "c"])       # all cached: the trunk is not called
    assert calls["n"] == 0 and torch.equal(v1, v2) and torch.equal(r1, r2)
    v3, _ = fe(t_img[[2, 0]], roi_img[[2, 0]], keys=["c", "d"])
    assert calls["n"] == 1 and torch.equal(v3[0], v1[2])
    img_net.conv1.forward = orig
    ft = VF.VisualFrontEnd(img_net, roi_net, if_fine_tune=True)
    v, r = ft(t_img, roi_img)
    (v.sum() + r.sum()).backward()
    assert img_net.conv1.weight.grad is not None and roi_net.layer4[0].weight.grad is not None
