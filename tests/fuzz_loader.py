"""Damaged model files through the host loader (psim_model_load_text + psim_model_prepare): truncated, characters replaced,
tokens inserted, ranges deleted, keys dropped or retyped.  The loader must answer with an error code or a model - never
crash (the reference catches nlohmann/json exceptions in InputManager::deserialize, inputManager.cpp:104-110).
Run by tests/test_host.py in a subprocess: usage  python tests/fuzz_loader.py <seed> <iterations>"""
import json, random, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psim_b200 import configs, lib as psim
base = json.dumps(configs.linear(num_cells=3, num_phonons=1000).to_dict())
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ok = bad = 0
alphabet = '{}[]",:0123456789.eE+-truefalsn \\u00e9\n\t'
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3000):
    s = list(base)
    kind = rng.randrange(5)
    if kind == 0:
        s = s[:rng.randrange(len(s))]
    elif kind == 1:
        for _ in range(rng.randrange(1, 4)):
            s[rng.randrange(len(s))] = rng.choice(alphabet)
    elif kind == 2:
        i = rng.randrange(len(s)); s[i:i] = list(rng.choice(['{', '[', '"', '\\', '1e999', '-', 'null', '\\u12', '[' * 600]))
    elif kind == 3:
        i = rng.randrange(len(s)); j = min(len(s), i + rng.randrange(1, 40)); del s[i:j]
    else:
        d = json.loads(base)
        # structural damage: drop / retype a random key somewhere
        def walk(o, path=()):
            out = [(o, path)]
            if isinstance(o, dict):
                for k, v in o.items(): out += walk(v, path + (k,))
            elif isinstance(o, list):
                for i, v in enumerate(o): out += walk(v, path + (i,))
            return out
        nodes = [n for n in walk(d) if isinstance(n[0], (dict, list)) and len(n[0])]
        o, _ = rng.choice(nodes)
        if isinstance(o, dict):
            k = rng.choice(list(o)); 
            if rng.random() < 0.5: del o[k]
            else: o[k] = rng.choice([None, "x", -1, 1e308, [], {}, True, 0])
        else:
            i = rng.randrange(len(o)); o[i] = rng.choice([None, "x", -1, [], {}, 0])
        s = list(json.dumps(d))
    text = ''.join(s)
    try:
        m = psim.Model(text=text)
        try:
            m.prepare()
        except psim.PsimError:
            pass
        m.close()
        ok += 1
    except psim.PsimError:
        bad += 1
    except UnicodeEncodeError:
        pass
print("accepted", ok, "rejected", bad)
