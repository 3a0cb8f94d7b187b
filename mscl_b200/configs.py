"""Programmatic builders for the model dicts of the MSCL pre-training configs.

The acceptance config is the reference's own file
(configs/recognition/moco/mscl_r18_cosm_lr2e-2.py), loaded unchanged by
`mscl_b200.Config.fromfile`.  That file does not travel to the GPU box, so bench.py and the
GPU tests build the same `model` dict here; tests/test_config_builds.py asserts the two are
equal whenever the reference tree is present.
"""


def _ce():
    return dict(type="CrossEntropyLoss_torch", ignore_index=-1)


def _moco(backbone, neck, basename, dim_in, K, max_iters, T):
    return dict(type="MoCoV2", backbone=dict(type=backbone), neck=neck,
                moco_head=dict(type="MoCoHead", basename=basename, loss_cls=_ce()),
                im_key="imgs", dim_in=dim_in, dim=128, K=K, m_base=0.994, max_iters=max_iters, T=T, mlp=True,
                aux_info=[], aug=dict(type="IdentityAug"))


def mscl_r18_model(K=65536, num_frames=8, image_size=112, dataset_size=219136, total_epochs=400, T=0.07,
                   aug="SyncMoCoAugmentV5"):
    """`model` of mscl_r18_cosm_lr2e-2.py: R3D-18 RGB branch with a TPN neck, slim r2d_18 flow branch."""
    max_iters = dataset_size * total_epochs
    tpn = dict(type="TPNMoCo", in_channels=[128, 256, 512], out_channels=128,
               sepc_cfg=dict(in_channels=[128, 128, 128], out_channels=128, stride=(2, 2, 2), iBN=False, Pconv_num=2))
    if aug == "SyncMoCoAugmentV5":
        aug_cfg = dict(type="SyncMoCoAugmentV5", crop_size=image_size, sync_level=("batch", "batch"),
                       t=(num_frames, num_frames), flow_suffix="flow_imgs", weak_aug=(False, False), visualize=True)
    else:
        aug_cfg = dict(type=aug)
    return dict(
        type="MSCLWithAug",
        recognizer=_moco("torchvision.r3d_18", tpn, "", 512, K, max_iters, T),
        recognizer_flow=_moco("resnet_flow.r2d_18", dict(type="BaseMoCo"), "flow", 128, K, max_iters, T),
        moco_mx_head=dict(type="MSCLWithAugMxHead", basename="mx", loss_cls=_ce(), same_kn=True, T=T),
        sup_head=dict(type="MSCLWithAugPosHeadV2", basename="", loss_pos=_ce(), bkb_channels=(None, None),
                      t=num_frames // 2, T=T,
                      aux_keys=dict(im_features=dict(q_mlvl="q_mlvl"), base_flow_features=dict(q_mlvl="q_flow_mlvl"),
                                    aug_flow_features=dict(q_mlvl="q_aug_flow_mlvl"))),
        im_key="imgs", flow_key="flow_imgs", aux_info=[], update_aug_flow=False, weight_aug_flow=(1.0, 1.0),
        aug=aug_cfg, same_kn=True)


def mscl_r50_model(K=65536, num_frames=8, image_size=224, dataset_size=219136, total_epochs=200, T=0.07,
                   aug="SyncMoCoAugmentV5"):
    """`model` of mscl_r50_cosm_lr3e-2.py: SlowOnly-R50 RGB branch with a TPN neck (one pyramid convolution), slim
    r2d_50 flow branch -- the large-parameter path of BASELINE config 5."""
    max_iters = dataset_size * total_epochs
    backbone = dict(type="ResNet3dSlowOnly", depth=50, pretrained=None, pretrained2d=False, lateral=False, num_stages=4,
                    conv1_kernel=(5, 7, 7), conv1_stride_t=2, pool1_stride_t=1, spatial_strides=(1, 2, 2, 2),
                    out_indices=(0, 1, 2, 3))
    tpn = dict(type="TPNMoCo", in_channels=[512, 1024, 2048], out_channels=128,
               sepc_cfg=dict(in_channels=[128, 128, 128], out_channels=128, stride=(1, 2, 2), iBN=False, Pconv_num=1))
    if aug == "SyncMoCoAugmentV5":
        aug_cfg = dict(type="SyncMoCoAugmentV5", crop_size=image_size, sync_level=("batch", "batch"),
                       t=(num_frames, num_frames), flow_suffix="flow_imgs", weak_aug=(False, False), visualize=True)
    else:
        aug_cfg = dict(type=aug)
    rgb = dict(type="MoCoV2", backbone=backbone, neck=tpn,
               moco_head=dict(type="MoCoHead", basename="", loss_cls=_ce()), im_key="imgs", dim_in=2048, dim=128, K=K,
               m_base=0.994, max_iters=max_iters, T=T, mlp=True, aux_info=[], aug=dict(type="IdentityAug"))
    flow = dict(type="MoCoV2", backbone=dict(type="resnet_flow.r2d_50"), neck=dict(type="BaseMoCo"),
                moco_head=dict(type="MoCoHead", basename="flow", loss_cls=_ce(), in_channels=256), im_key="imgs",
                dim_in=256, dim=128, K=K, m_base=0.994, max_iters=max_iters, T=T, mlp=True, aux_info=[],
                aug=dict(type="IdentityAug"))
    return dict(
        type="MSCLWithAug", recognizer=rgb, recognizer_flow=flow,
        moco_mx_head=dict(type="MSCLWithAugMxHead", basename="mx", loss_cls=_ce(), same_kn=True, T=T),
        sup_head=dict(type="MSCLWithAugPosHeadV2", basename="", loss_pos=_ce(), bkb_channels=(None, 256),
                      t=num_frames // 2, T=T,
                      aux_keys=dict(im_features=dict(q_mlvl="q_mlvl"), base_flow_features=dict(q_mlvl="q_flow_mlvl"),
                                    aug_flow_features=dict(q_mlvl="q_aug_flow_mlvl"))),
        im_key="imgs", flow_key="flow_imgs", aux_info=[], update_aug_flow=False, weight_aug_flow=(1.0, 1.0),
        aug=aug_cfg, same_kn=True)
