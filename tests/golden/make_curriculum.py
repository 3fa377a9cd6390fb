"""Generates myochallenge_b200/assets/curriculum/baoding_winner.json from the reference's curriculum directory
(/root/reference/trained_models/curriculum_steps_complete_baoding_winner/NN_*/config.json: the env kwargs of each of the 32
training steps, plain data) plus, from each step's main.py, the env name it trains on. Run in the build container:
    python tests/golden/make_curriculum.py"""
import json
import os
import re

REF = "/root/reference/trained_models/curriculum_steps_complete_baoding_winner"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "myochallenge_b200", "assets", "curriculum",
                   "baoding_winner.json")


def main():
    steps = []
    for d in sorted(os.listdir(REF)):
        cfg_path = os.path.join(REF, d, "config.json")
        if not os.path.exists(cfg_path):
            continue
        cfg = json.load(open(cfg_path))
        env_name = None
        main_py = os.path.join(REF, d, "main.py")
        if os.path.exists(main_py):
            m = re.search(r'^\s*env_name\s*=\s*["\']([^"\']+)["\']', open(main_py).read(), re.M)
            env_name = m.group(1) if m else None
        steps.append({"step": d, "env_name": env_name, "config": cfg})
    json.dump({"source": "trained_models/curriculum_steps_complete_baoding_winner/*/config.json of amathislab/myochallenge", "steps": steps},
              open(OUT, "w"), indent=1)
    print("wrote", OUT, len(steps), "steps; env names:", sorted({s["env_name"] for s in steps}, key=str))


if __name__ == "__main__":
    main()
