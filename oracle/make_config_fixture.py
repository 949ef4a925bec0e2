"""Test infrastructure (not product): parses the reference's real YAML configs with the reference's own
configs/parser.py (YAMLParser + combine_entries, exactly what train_flow_parallel_supervised_SNN.py:45,537 do) and writes
the resulting dicts as JSON fixtures, so that the GPU box (which has no /root/reference) can replay the scripts' model
construction with the shipped configuration values.

    python oracle/make_config_fixture.py        # writes tests/golden/ref_config_*.json
"""
import json
import os
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CONFIGS = {
    "ref_config_dsec_en4.json": "configs/train_DSEC_supervised_SDformerFlow_en4.yml",
    "ref_config_mdr.json": "configs/train_MDR_supervised_SDformerFlow.yml",
}


def main():
    sys.path.insert(0, REF)
    from configs.parser import YAMLParser
    for out, yml in CONFIGS.items():
        p = YAMLParser(os.path.join(REF, yml))
        cfg = p.combine_entries(p.config)
        with open(os.path.join(OUT, out), "w") as f:
            json.dump({"source": yml, "config": cfg}, f, indent=1, sort_keys=True)
        print(out, cfg["model"]["name"], cfg["model"]["spiking_neuron"]["neuron_type"])


if __name__ == "__main__":
    main()
