"""Checkpoint interchange with the REFERENCE's own writer / reader, both directions, strict=True.

Runs in the build container only (it needs /root/reference; skipped on the GPU box): the reference's
`dump_models_into_disk` (main.py:105-177) and `load_clean_weights` (main.py:54-70) are compiled from the reference source at
run time by tests/golden/make_golden.py::reference_checkpoint_functions, reference modules come from /root/reference/models.
The GPU-side half (a reference-written file loaded into the drop-in on the device) is tests/test_valpost_gpu.py-style
fixture based: tests/test_parity_gpu.py::test_reference_written_checkpoint_loads_into_dropin.
"""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("JMT_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="needs the reference checkout")


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    R = mg._import_reference()
    dump, load = mg.reference_checkpoint_functions()
    return R, dump, load


def _same(sd_a, sd_b):
    assert list(sd_a.keys()) == list(sd_b.keys())
    for k in sd_a:
        assert torch.equal(sd_a[k].cpu(), sd_b[k].cpu()), k


def test_reference_written_checkpoints_load_into_dropins(ref, tmp_path):
    """reference modules -> reference dump_models_into_disk -> jmt_b200.checkpoint.load_models_from_disk (strict=True)."""
    import jmt_b200
    from jmt_b200 import checkpoint as CK
    R, dump, _ = ref
    torch.manual_seed(5)
    fusion = R["Two_transformers"](0.0, 0.0, 2, 1, "TRANSFORMER", "SELF_ATTEN", 512)
    pre = R["SingleBackbonePretrainer"](0.0, 0.0)
    fca = R["FcLayer"](768, 512)
    tra = R["Intra"](512, 1, 512, 1)
    dump({}, fusion, pre, fca, tra, None, None, 1, str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == sorted(["fusion_w.pt", "backbone_pretrainer_w.pt", "fc_layer_for_audio_concat.pt",
                                                   "transformer_audio_modality_fusion.pt"])
    mine = {"fusion_model": jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "SELF_ATTEN", 512),
            "backbone_pretrainer": jmt_b200.SingleBackbonePretrainer(0.0, 0.0),
            "fc_layer_for_audio_concat": jmt_b200.FcLayer(768, 512),
            "transformer_audio_modality_fusion": jmt_b200.Intra_modal_transformer_fusion(512, 1, 512, 1)}
    CK.load_models_from_disk(str(tmp_path), mine)
    for role, r in (("fusion_model", fusion), ("backbone_pretrainer", pre), ("fc_layer_for_audio_concat", fca),
                    ("transformer_audio_modality_fusion", tra)):
        _same(mine[role].state_dict(), r.state_dict())


def test_dropin_written_checkpoints_load_into_reference(ref, tmp_path):
    """drop-in modules -> jmt_b200.checkpoint.dump_models_into_disk -> reference load_clean_weights + load_state_dict(strict=True)
    (main.py:515-518), incl. a DataParallel-prefixed file."""
    import jmt_b200
    from jmt_b200 import checkpoint as CK
    R, _, load = ref
    torch.manual_seed(6)
    fusion = jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "NONE", "FC", 512)
    fcv = jmt_b200.FcLayer(1280, 512)
    CK.dump_models_into_disk(str(tmp_path), {"fusion_model": fusion, "fc_layer_for_video_concat": fcv}, epoch=2)
    rf = R["Two_transformers"](0.0, 0.0, 1, 1, "NONE", "FC", 512)
    rf.load_state_dict(load(os.path.join(tmp_path, "fusion_w.pt"), map_location="cpu"), strict=True)
    _same(rf.state_dict(), fusion.state_dict())
    rc = R["FcLayer"](1280, 512)
    sd = torch.load(os.path.join(tmp_path, "fc_layer_for_video_concat.pt"), weights_only=True)
    torch.save({"module." + k: v for k, v in sd.items()}, os.path.join(tmp_path, "dp.pt"))
    rc.load_state_dict(load(os.path.join(tmp_path, "dp.pt"), map_location="cpu"), strict=True)
    _same(rc.state_dict(), fcv.state_dict())


def test_tcn_head_out_of_a_reference_i3d_checkpoint(ref, tmp_path):
    """The reference keeps the TCN inside I3D_WSDDA (`temporal.*` keys of vision_i3d.pt, I3DWSDDA.py:26-28; main.py:157-159):
    checkpoint.extract_submodule_state pulls it out for the drop-in TemporalConvNet (strict=True, incl. the net.* aliases)."""
    import jmt_b200
    from jmt_b200 import checkpoint as CK
    R, _, load = ref
    torch.manual_seed(7)
    tcn = R["TCN"](1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1)
    sd = {"temporal." + k: v for k, v in tcn.state_dict().items()}
    sd["i3d_WSDDA.logits.conv3d.weight"] = torch.zeros(2, 2)            # unrelated backbone tensors next to it
    torch.save(sd, os.path.join(tmp_path, "vision_i3d.pt"))
    mine = jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1)
    mine.load_state_dict(CK.extract_submodule_state(CK.load_clean_weights(os.path.join(tmp_path, "vision_i3d.pt")), "temporal."),
                         strict=True)
    _same(mine.state_dict(), tcn.state_dict())
    with pytest.raises(KeyError):
        CK.extract_submodule_state(sd, "nothing.")
