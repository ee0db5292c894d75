"""CPU: the drop-in classes keep the reference's constructor signatures, module tree, state_dict keys and
parameter counts (SURVEY.md §4 pins, §8b), are picklable, refuse CPU tensors and support shape probing."""
import io
import pickle

import pytest
import torch
import torch.nn as nn

import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import models as M, workloads as W


def test_param_counts(golden):
    assert sum(p.numel() for p in W.build_c1(M).parameters()) == golden["param_counts"]["c1"] == 11226882
    assert sum(p.numel() for p in W.build_c2(M).parameters()) == golden["param_counts"]["c2_gru"] == 1707778
    assert sum(p.numel() for p in W.build_c3(M).parameters()) == golden["param_counts"]["c3"] == 19697668
    assert sum(p.numel() for p in W.build_c2(M, heads=("LSTM_1L",)).parameters()) == 2233090
    assert sum(p.numel() for p in W.build_c2(M, heads=("Avg_features",)).parameters()) == 131842


def test_module_tree_equals_reference_printout(golden):
    torch.manual_seed(0)
    assert str(W.build_c3(M)) == golden["c3_module_tree"]


def test_module_trees_of_the_other_assemblies_equal_the_reference(golden):
    """golden_v2: text branch, averaged fusion, base PhysVerbClassifier, the older MultimodalModel, AudioTextualModel —
    the drop-in classes print the same module tree as the reference's (same sub-module names => same state_dict keys),
    and the grad_norms recorded from the reference name exactly the drop-in's parameters."""
    for name, tree in golden["module_trees"].items():
        spec = golden["cases"][name]["spec"]
        torch.manual_seed(0)
        model = getattr(W, spec["builder"])(M, **spec["bkw"])
        assert str(model) == tree, name
        assert set(dict(model.named_parameters())) == set(golden["cases"][name]["grad_norms"]), name


def test_state_dict_keys():
    sd = W.build_c3(M).state_dict()
    for k in ("modality_extractors_dict.audio.transformer_squence_processing.layers.0.self_attn.in_proj_weight",
              "modality_extractors_dict.video.feature_extractor.embedding.0.weight",
              "modality_fusion_module.modality_fusion_transformer.norm.bias",
              "classifiers.adaptors_dict.audio.0.weight", "classifiers.classifiers_dict.phys.3.bias"):
        assert k in sd
    sd2 = W.build_c2(M).state_dict()
    assert set(sd2) == {f"models_dict.GRU_1L.{s}" for s in (
        "sequence_nn.weight_ih_l0", "sequence_nn.weight_hh_l0", "sequence_nn.bias_ih_l0", "sequence_nn.bias_hh_l0",
        "output_classifier.0.weight", "output_classifier.0.bias", "output_classifier.3.weight", "output_classifier.3.bias")}
    sd1 = W.build_c1(M).state_dict()
    assert "1.classifier.1.weight" in sd1 and "1.classifier.4.bias" in sd1


def test_models_are_picklable():
    model = W.build_c3(M)
    blob = pickle.dumps(model)
    again = pickle.loads(blob)
    assert str(again) == str(model)
    buf = io.BytesIO()
    torch.save(model, buf)


def test_cpu_tensor_is_refused_loudly():
    model = W.build_c1(M)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.zeros(2, 10, 768))


def test_shape_probe_on_cpu():
    """train_multimodal.py:346-353 probes feature shapes by running the extractors on CPU zeros."""
    model = W.build_c3(M, t_audio=20, t_video=8)
    with mar.shape_probe():
        a = model.modality_extractors_dict["audio"](torch.zeros(1, 20, 768))
        v = model.modality_extractors_dict["video"](torch.zeros(1, 8, 512))
        assert a.shape == (1, 20, 768) and v.shape == (1, 8, 768)
        data, _ = W.batch_c3(B=2, t_audio=20, t_video=8)
        out = model(data)
        assert out["phys"].shape == (2, 2) and out["verb"].shape == (2, 2)
        out2 = W.build_c2(M, heads=("GRU_1L", "LSTM_1L", "Avg_features"))(torch.zeros(3, 5, 512))
        assert all(v.shape == (3, 2) for v in out2.values())


def test_get_output_names_and_models_names():
    assert W.build_c3(M).get_output_names() == ["phys", "verb"]
    assert W.build_c2(M, heads=("LSTM_1L", "GRU_1L")).get_models_names() == ["LSTM_1L", "GRU_1L"]


def test_losses_dict_contract():
    ld = M.LossesDict()
    assert hasattr(ld, "backward") and isinstance(ld, dict)
    ld.backward()     # empty: no-op, like the reference when no head is active


def test_precision_switch():
    assert mar.get_precision() == torch.bfloat16
    with mar.precision("fp32"):
        assert mar.get_precision() == torch.float32
    assert mar.get_precision() == torch.bfloat16
    with pytest.raises(ValueError):
        mar.set_precision("fp16")


def test_unsupported_layer_config_is_refused():
    enc = nn.TransformerEncoder(nn.TransformerEncoderLayer(64, 4, batch_first=True, norm_first=True), 1)
    with pytest.raises(NotImplementedError):
        M._check_layer(enc.layers[0])


def test_audio_multi_nn_structure():
    heads = W.build_c2(M, heads=("GRU_1L", "LSTM_1L")).models_dict
    m = M.AudioMultiNN({k: v for k, v in heads.items()}, {"w2v": nn.Identity()})
    assert m.get_models_names() == (["w2v"], ["GRU_1L", "LSTM_1L"])
    assert not m.extractor_dict.training                      # the reference freezes the extractor at construction
    assert str(pickle.loads(pickle.dumps(m))) == str(m)
