"""Drop-in for the CLI of the reference's feature_extraction/run_all.py (`--input --segmentation --output`,
:392-520) for the steps that sit on the voxel hot path: step 3 (lesion multiplicity) and step 4 (morphology and
margins) run on the device and write `step3_multiplicity.json` / `step4_morphology.json` with the reference's keys.

Steps 1, 2, 5, 6 (intensity heuristics, mass effect, quality control, normal structures) and the narrative report /
LLM summary compiled from all six are outside this package's scope: `comprehensive_analysis.json`
lists them under "steps_not_run" instead of carrying their sections.
"""
import argparse
from datetime import datetime
from pathlib import Path

from . import utils as U
from .step3_multiplicity import analyze_multiplicity
from .step4_morphology import analyze_morphology

STEPS_NOT_RUN = ("step1_sequence_findings", "step2_mass_effect", "step5_quality", "step6_normal_structures")
_RULE = "=" * 70


def run_all_steps(input_folder, segmentation_path, output_folder):
    source, target = Path(input_folder), Path(output_folder)
    target.mkdir(parents=True, exist_ok=True)
    case_id = U.get_case_id(source)
    print(f"{_RULE}\nBRAIN MRI FEATURE EXTRACTION PIPELINE\n{_RULE}")
    print(f"\nCase ID: {case_id}\nInput: {source}\nSegmentation: {segmentation_path}\nOutput: {target}\n")
    sections = {}
    for key, title, driver in (("step3_multiplicity", "RUNNING STEP 3: Lesion Multiplicity", analyze_multiplicity),
                               ("step4_morphology", "RUNNING STEP 4: Tumor Morphology", analyze_morphology)):
        print(f"\n{_RULE}\n{title}\n{_RULE}")
        sections[key] = driver(source, segmentation_path, target / f"{key}.json")
    everything = {"case_id": case_id, "analysis_timestamp": datetime.now().isoformat(), "input_folder": str(source),
                  "segmentation_path": str(segmentation_path), **sections, "steps_not_run": list(STEPS_NOT_RUN)}
    combined = target / "comprehensive_analysis.json"
    U.save_results(everything, combined)
    print(f"\n{_RULE}\nANALYSIS COMPLETE\n{_RULE}\n\nOutput files:\n  • Comprehensive analysis: {combined}\n"
          f"  • Individual step results: step3 / step4 JSON files")
    return everything


def main(argv=None):
    cli = argparse.ArgumentParser(description="Run the device-side feature extraction steps (3 and 4)")
    cli.add_argument("--input", required=True, help="Input folder containing MRI sequences")
    cli.add_argument("--segmentation", required=True, help="Path to segmentation mask (NIfTI)")
    cli.add_argument("--output", required=True, help="Output folder for all results")
    args = cli.parse_args(argv)
    run_all_steps(args.input, args.segmentation, args.output)


if __name__ == "__main__":
    main()
