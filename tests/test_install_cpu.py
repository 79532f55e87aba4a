"""install() patches the replacements into the reference tree without touching gens.py / runner.py.
Needs /root/reference (build container only); skipped on the GPU box."""
import os
import sys
import types

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_install_patches_reference_modules():
    sys.path.insert(0, REF)
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))
    try:
        import gens_b200
        from gens_b200.config import gens_model_conf
        mods = gens_b200.install()
        assert mods["volume"].Volume is gens_b200.Volume
        assert mods["implicit_surface"].ImplicitSurface is gens_b200.ImplicitSurface
        import models.modules.sdf_network as ref_sdf
        assert ref_sdf.lookup_volume is gens_b200.lookup_volume
        import models.modules.reg_network as ref_reg
        assert ref_reg.RegNetwork is gens_b200.RegNetwork
        assert set(ref_reg.RegNetwork(gens_model_conf()["reg_network"]).state_dict()) >= {
            "conv0.conv.weight", "encoder_layers.4.1.conv.weight", "decoder_layers.0.conv.weight", "out_layers.2.bias"}
        # the reference's own JIT build of gridsample_grad2 must not have been triggered
        assert "gridsample_grad2" not in sys.modules
        # constructor / method surface the callers rely on (gens.py:70, :143, :155)
        from gens_b200.config import gens_model_conf
        conf = gens_model_conf()
        vol = mods["volume"].Volume(conf["volume"])
        assert vol.volume_dims == [256, 128, 64, 32, 16] and hasattr(vol, "agg_mean_var")
        surf = mods["implicit_surface"].ImplicitSurface(conf["implicit_surface"])
        for name in ("up_sample", "cat_z_vals", "render_core", "render", "extract_geometry", "validate", "forward",
                     "tv_regularization"):
            assert callable(getattr(surf, name))
        keys = set(surf.state_dict().keys())
        assert {"sdf_network.lin0.weight_g", "sdf_network.lin6.weight_v", "deviation_network.variance",
                "color_network.s", "color_network.rgb_fc.4.bias"} <= keys
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_install_can_leave_the_regulariser_alone():
    sys.path.insert(0, REF)
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))
    try:
        import gens_b200
        gens_b200.install(regulariser=False)
        import models.modules.reg_network as ref_reg
        assert ref_reg.RegNetwork is not gens_b200.RegNetwork and ref_reg.RegNetwork.__module__ == "models.modules.reg_network"
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
