"""ARaymarchVolume's per-tick update policy (RaymarchVolume.cpp:326-465) mirrored in raymarch_volume.py, checked on the CPU
with a recording stand-in for the operator surface: which of Clear / AddDirLight / ChangeDirLight runs, with which arguments."""
import pytest

from tbraymarcherplugin_b200.raymarch_utils import FBasicRaymarchRenderingResources, FTransform
from tbraymarcherplugin_b200.raymarch_volume import ARaymarchClipPlane, ARaymarchLight, ARaymarchVolume, ERaymarchMaterial


class RecordingOps:
    def __init__(self, fail_on=None):
        self.calls, self.fail_on = [], fail_on

    def ClearResourceLightVolumes(self, res, value):
        self.calls.append(("clear", value))

    def AddDirLightToSingleVolume(self, res, light, added, world, bGPUSync=False, stats=None):
        self.calls.append(("add", tuple(light.LightDirection), light.LightIntensity, added, bGPUSync))
        return self.fail_on != "add"

    def ChangeDirLightInSingleVolume(self, res, old, new, world, bGPUSync=False, stats=None):
        self.calls.append(("change", tuple(old.LightDirection), tuple(new.LightDirection), old.LightIntensity, new.LightIntensity))
        return self.fail_on != "change"


def make(n_lights=4, **kw):
    res = FBasicRaymarchRenderingResources()
    res.bIsInitialized = True
    lights = [ARaymarchLight((1.0, 0.1 * i, -0.3), 1.0, f"L{i}") for i in range(n_lights)]
    ops = RecordingOps(kw.pop("fail_on", None))
    return ARaymarchVolume(res, lights, ops=ops, **kw), lights, ops


def test_nothing_changed_means_no_gpu_work_and_uninitialised_volumes_do_not_tick():
    vol, lights, ops = make()
    assert vol.Tick().action == "none" and ops.calls == []
    vol.RaymarchResources.bIsInitialized = False
    lights[0].LightIntensity = 0.5
    assert vol.Tick().action == "not_initialized" and ops.calls == []


def test_one_changed_light_is_an_incremental_change_with_the_remembered_old_parameters():
    vol, lights, ops = make()
    lights[2].ForwardVector = (0.0, 1.0, 0.0)
    rep = vol.Tick()
    assert rep.action == "incremental" and rep.lights_updated == 1
    assert ops.calls == [("change", (1.0, 0.2, -0.3), (0.0, 1.0, 0.0), 1.0, 1.0)]
    assert vol.Tick().action == "none"  # the map was refreshed (RaymarchVolume.cpp:411)


def test_reset_rule_more_than_one_and_at_least_half_of_the_lights():
    vol, lights, ops = make(4)
    lights[0].LightIntensity = 0.2
    lights[1].LightIntensity = 0.3
    rep = vol.Tick()  # 2 changed, 2 >= 4 // 2 -> full reset
    assert rep.action == "reset" and [c[0] for c in ops.calls] == ["clear", "add", "add", "add", "add"]
    assert all(c[4] is True for c in ops.calls[1:]), "bFastShader selects the fused sweep"
    # with 6 lights, 2 changed lights are below half: two incremental changes
    vol, lights, ops = make(6)
    lights[0].LightIntensity = 0.2
    lights[5].LightIntensity = 0.3
    assert vol.Tick().action == "incremental" and [c[0] for c in ops.calls] == ["change", "change"]
    # a single light never triggers the rule ("Num() > 1"), even if it is the only one
    vol, lights, ops = make(1)
    lights[0].LightIntensity = 0.2
    assert vol.Tick().action == "incremental"


def test_world_or_clip_plane_change_forces_a_full_reset_with_the_new_world():
    plane = ARaymarchClipPlane((0, 0, 0), (0, 0, 1))
    vol, lights, ops = make(3, ClippingPlane=plane)
    vol.ComponentTransform = FTransform((5.0, 0.0, 0.0))
    assert vol.Tick().action == "reset" and vol.bRequestedRecompute is False
    ops.calls.clear()
    vol.ComponentTransform = FTransform((5.0 + 5e-5, 0.0, 0.0))  # within FTransform::Equals' tolerance: no change
    assert vol.Tick().action == "none"
    plane.Center = (0.0, 0.0, 0.1)  # clip planes compare exactly
    assert vol.Tick().action == "reset"
    # only the Lit material keeps the light volume up to date (RaymarchVolume.cpp:366-368)
    vol.SelectRaymarchMaterial = ERaymarchMaterial.Intensity
    ops.calls.clear()
    lights[0].LightIntensity = 0.1
    assert vol.Tick().action == "none" and ops.calls == []


def test_new_lights_enter_the_map_and_a_failed_reset_is_retried():
    vol, lights, ops = make(4)
    extra = ARaymarchLight((0.0, 0.0, -1.0), 0.7, "late")
    vol.LightsArray.append(extra)
    rep = vol.Tick()  # unknown light: recorded, then "changed" from its own current parameters
    assert rep.action == "incremental" and ops.calls[0][0] == "change" and ops.calls[0][1] == ops.calls[0][2]
    vol2, _, ops2 = make(2, fail_on="add")
    vol2.bRequestedRecompute = True
    rep = vol2.Tick()
    assert rep.errors and vol2.bRequestedRecompute is True  # RaymarchVolume.cpp:441-446 returns before clearing the flag


def test_reference_quirk_stale_light_map_after_a_reset_and_the_opt_in_fix():
    for refresh in (False, True):
        vol, lights, ops = make(4, bRefreshLightMapOnReset=refresh)
        lights[0].LightIntensity = 0.2
        vol.bRequestedRecompute = True
        assert vol.Tick().action == "reset"
        ops.calls.clear()
        rep = vol.Tick()
        if refresh:
            assert rep.action == "none"
        else:  # the reference re-issues a change from parameters the light volume no longer holds
            assert rep.action == "incremental" and ops.calls[0][:1] == ("change",) and ops.calls[0][3:] == (1.0, 0.2)


class RecordingMaterialOps(RecordingOps):
    def GenerateOctree(self, res):
        self.calls.append(("octree",))

    def PerformWindowedLitRaymarch(self, res, cam, world, steps, rows=None):
        self.calls.append(("lit", steps))
        return None, 0

    def PerformWindowedIntensityRaymarch(self, res, cam, world, steps, rows=None):
        self.calls.append(("intensity", steps))
        return None, 0

    def PerformWindowedRaymarchOctree(self, res, cam, world, steps, mip, rows=None):
        self.calls.append(("octree_march", steps, mip))
        return None, 0


def test_octree_is_rebuilt_once_and_only_while_the_octree_material_is_selected():
    """RaymarchVolume.cpp:358-363 + :553-554: a new volume requests the rebuild; the tick performs it only under the octree material,
    and lights are only maintained under the lit material (:365-368)."""
    res = FBasicRaymarchRenderingResources()
    res.bIsInitialized = True
    ops = RecordingMaterialOps()
    lights = [ARaymarchLight((1.0, 0.0, -0.3), 1.0, "L0")]
    vol = ARaymarchVolume(res, lights, ops=ops)
    vol.OnVolumeLoaded()
    rep = vol.Tick()  # lit material: full reset (bRequestedRecompute), the octree request stays pending
    assert rep.action == "reset" and not rep.octree_rebuilt and ("octree",) not in ops.calls and vol.bRequestedOctreeRebuild
    ops.calls.clear()
    vol.SelectRaymarchMaterial = ERaymarchMaterial.Octree
    lights[0].ForwardVector = (0.0, 1.0, 0.0)
    rep = vol.Tick()
    assert rep.octree_rebuilt and ops.calls == [("octree",)] and not vol.bRequestedOctreeRebuild  # no light work under this material
    assert vol.Tick().octree_rebuilt is False and ops.calls == [("octree",)]
    vol.OctreeVolumeMip, vol.RaymarchingSteps = 2, 200.0
    vol.Render(camera := object())
    vol.SelectRaymarchMaterial = ERaymarchMaterial.Intensity
    vol.Render(camera)
    vol.SelectRaymarchMaterial = ERaymarchMaterial.Lit
    vol.Render(camera)
    assert ops.calls[1:] == [("octree_march", 200.0, 2), ("intensity", 200.0), ("lit", 200.0)]
    assert vol.Tick().action == "incremental"  # back under the lit material the moved light is finally updated


def test_set_volume_asset_scales_the_mesh_and_requests_recompute_and_octree_rebuild():
    """RaymarchVolume.cpp:467-560: TF (default ramp without a curve), windowing, mesh scale = WorldDimensions / 10 (mm -> cm), full light
    recompute and octree rebuild requested."""
    class Ops(RecordingMaterialOps):
        def MakeDefaultTFTexture(self, res):
            self.calls.append(("default_tf",))

        def ColorCurveToTexture(self, res, curve):
            self.calls.append(("curve_tf",))

        def SetWindowingParameters(self, res, w):
            self.calls.append(("window", w.Center, w.Width))

    class Info:
        WorldDimensions = (320.0, 240.0, 125.0)

    from tbraymarcherplugin_b200.raymarch_utils import FWindowingParameters

    ops = Ops()
    vol = ARaymarchVolume(FBasicRaymarchRenderingResources(), [ARaymarchLight((1.0, 0.0, -0.3), 1.0, "L0")], ops=ops)
    assert vol.SetVolumeAsset(FBasicRaymarchRenderingResources()) is False  # no texture: nothing happens (:469-472)
    res = FBasicRaymarchRenderingResources()
    res._h = object()  # stands for a created resource set
    assert vol.SetVolumeAsset(res, Info(), None, FWindowingParameters(0.4, 0.3, True, False))
    assert ops.calls == [("default_tf",), ("window", 0.4, 0.3)]
    assert vol.ComponentTransform.Scale3D == (32.0, 24.0, 12.5) and vol.WorldParameters.VolumeTransform.Scale3D == (32.0, 24.0, 12.5)
    assert vol.bRequestedRecompute and vol.bRequestedOctreeRebuild and res.bIsInitialized
    ops.calls.clear()
    assert vol.Tick().action == "reset" and ops.calls[0] == ("clear", 0.0)
    res._h = None  # do not let __del__ hand the stand-in to the library


def test_setters_request_a_recompute_only_when_something_changed():
    """RaymarchVolume.cpp:562-577, 746-818: window centre / width / cut-offs and the TF curve invalidate the light volume; unchanged values
    do nothing; steps and renderer switches do not touch the lights."""
    class Ops(RecordingMaterialOps):
        def SetWindowingParameters(self, res, w):
            self.calls.append(("window", w.Center, w.Width, w.LowCutoff, w.HighCutoff))

        def ColorCurveToTexture(self, res, curve):
            self.calls.append(("curve_tf",))

        def FlushRenderingCommands(self, res):
            self.calls.append(("flush",))

    res = FBasicRaymarchRenderingResources()
    res.bIsInitialized = True
    ops = Ops()
    vol = ARaymarchVolume(res, [ARaymarchLight((1.0, 0.0, -0.3), 1.0, "L0")], ops=ops)
    vol.SetWindowCenter(vol.GetWindowCenter())
    vol.SetLowCutoff(True)
    assert ops.calls == [] and not vol.bRequestedRecompute
    vol.SetWindowCenter(0.4)
    assert ops.calls == [("window", 0.4, 1.0, True, True)] and vol.bRequestedRecompute
    assert vol.Tick().action == "reset" and not vol.bRequestedRecompute
    ops.calls.clear()
    vol.SetWindowWidth(0.5), vol.SetHighCutoff(False)
    assert [c[0] for c in ops.calls] == ["window", "window"] and ops.calls[-1] == ("window", 0.4, 0.5, True, False) and vol.bRequestedRecompute
    vol.Tick()
    ops.calls.clear()
    vol.SetTFCurve(None)
    assert ops.calls == [] and not vol.bRequestedRecompute
    vol.SetTFCurve(object())
    assert ops.calls == [("curve_tf",), ("flush",)] and vol.bRequestedRecompute
    vol.Tick()
    vol.SetRaymarchSteps(300), vol.SwitchRenderer(ERaymarchMaterial.Intensity)
    assert vol.RaymarchingSteps == 300.0 and vol.SelectRaymarchMaterial == ERaymarchMaterial.Intensity and not vol.bRequestedRecompute

    class Loader:  # a loader that cannot produce an asset: the reference returns false and keeps the old volume
        @staticmethod
        def CreateVolumeFromFile(*a, **k):
            raise RuntimeError("no such file")

    assert vol.LoadMHDFileIntoVolumeNormalized("missing.mhd", loader=Loader) is False and vol.RaymarchResources is res


def test_joined_reset_adds_all_lights_in_one_call():
    class Ops(RecordingOps):
        def AddDirLightsToSingleVolumeJoined(self, res, lights, added, world, stats=None):
            self.calls.append(("joined", len(lights), added))
            return self.fail_on != "joined"

    res = FBasicRaymarchRenderingResources()
    res.bIsInitialized = True
    lights = [ARaymarchLight((1.0, 0.1 * i, -0.3), 1.0, f"L{i}") for i in range(3)] + [None]
    ops = Ops()
    vol = ARaymarchVolume(res, lights, ops=ops)
    vol.bJoinSameAxisLights, vol.bRequestedRecompute = True, True
    rep = vol.Tick()
    assert rep.action == "reset" and rep.lights_updated == 3 and ops.calls == [("clear", 0.0), ("joined", 3, True)] and not vol.bRequestedRecompute
    ops.calls.clear()
    ops.fail_on, vol.bRequestedRecompute = "joined", True
    rep = vol.Tick()
    assert rep.errors and vol.bRequestedRecompute  # retried next tick, like the per-light path
