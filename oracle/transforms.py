"""Oracle (test infrastructure, never imported by the product): plain-Python restatement of the reference's
post-ensemble transform chain -- which stem a transform applies to, in which order the models run, and which of a
model's two outputs is kept (reference: /root/reference/modules/separator/stem_separator.py:679-700
_should_apply_transform, :777-839 _apply_transform_chain, :903-934 the call order inside predict_with_model).

`run_model(model_file, array) -> [(output name, array), ...]` stands for `separator.load_model` + `separator.separate`
(primary stem first, secondary second, names as they appear in the output file names)."""
from typing import Callable, Dict, List, Sequence, Tuple


def should_apply_transform(stem_name: str, setting: str) -> bool:
    # stem_separator.py:680-700
    if setting == "Nothing":
        return False
    if setting == "All":
        return True
    if setting == "All Vocals":
        return "vocals)" in stem_name.lower()
    if setting == "Main Vocals":
        return "vocals)" in stem_name and "(bg_vocals" not in stem_name.lower()
    return False


def transformations(opts: Dict) -> List[Tuple[str, str, str]]:
    # stem_separator.py:795-800 (the models of :147-149 by default)
    return [
        ("dereverb_mel_band_roformer_anvuew_sdr_19.1729.ckpt", "No Reverb", opts.get("reverb_removal", "Nothing")),
        (opts.get("delay_removal_model", "dereverb-echo_mel_band_roformer_sdr_13.4843_v2.ckpt"), "dry", opts.get("echo_removal", "Nothing")),
        (opts.get("crowd_removal_model", "UVR-MDX-NET_Crowd_HQ_1.onnx"), "No Crowd", opts.get("crowd_removal", "Nothing")),
        (opts.get("noise_removal_model", "UVR-DeNoise.pth"), "No Noise", opts.get("noise_removal", "Nothing")),
    ]


def apply_transform_chain(array, stem_label: str, opts: Dict, run_model: Callable, skip_transforms: Sequence[str] = ()):
    # stem_separator.py:777-839
    current = array
    simulated_name = f"({stem_label})"
    for model_file, out_label, flag in transformations(opts):
        if out_label in skip_transforms:
            continue
        if should_apply_transform(simulated_name, flag):
            outs = run_model(model_file, current)
            key = out_label.replace(" ", "").lower()
            chosen = None
            if len(outs) == 2:
                chosen = outs[0] if key in outs[0][0].replace(" ", "").lower() else outs[1]
            else:
                for o in outs:
                    if key in o[0].replace(" ", "").lower():
                        chosen = o
                        break
            if chosen is not None:
                current = chosen[1]
    return current


def transform_stage(results: Dict[str, Dict], opts: Dict, run_model: Callable) -> None:
    """The part of predict_with_model between the ensemble and the multi-stem stage (:903-934), without the background
    vocal split: reverb removal runs the WHOLE chain on the vocals first; when crowd or noise removal is set the chain runs
    again on the vocals without its reverb step, and on the instrumental."""
    if opts.get("reverb_removal", "Nothing") != "Nothing":
        for res in results.values():
            if res.get("vocals") is not None:
                res["vocals"] = apply_transform_chain(res["vocals"], "vocals", opts, run_model)
    if any(opts.get(k, "Nothing") != "Nothing" for k in ("crowd_removal", "noise_removal")):
        for res in results.values():
            if res.get("vocals") is not None:
                res["vocals"] = apply_transform_chain(res["vocals"], "vocals", opts, run_model, skip_transforms=["No Reverb"])
            if res.get("instrumental") is not None:
                res["instrumental"] = apply_transform_chain(res["instrumental"], "instrumental", opts, run_model)
