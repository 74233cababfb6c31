"""CPU: the plugin's constructors accept the model{} / loss{} blocks of every conf file the reference ships
(code/confs/{dtu,bmvs,abc-neat-a}.conf) -- parameter names and shapes as in the reference's own modules.  Needs the
reference tree (build container only); skipped elsewhere.  The few lines of HOCON the confs use are parsed here."""
import os
import re

import pytest
import torch

REF_CONFS = os.path.join(os.environ.get("NEAT_REFERENCE_ROOT", "/root/reference"), "code", "confs")
CONFS = ["dtu.conf", "bmvs.conf", "abc-neat-a.conf"]


def parse_hocon(text):
    """The subset the shipped confs use: nested `name { ... }` / `name{`, `key = value`, [lists], # comments."""
    text = re.sub(r"#.*", "", text)
    tokens = re.findall(r"\[[^\]]*\]|[{}=]|[^\s{}=\[\]]+", text)
    pos = 0

    def value(tok):
        if tok.startswith("["):
            return [value(t.strip()) for t in tok[1:-1].split(",") if t.strip()]
        if tok in ("True", "true"):
            return True
        if tok in ("False", "false"):
            return False
        try:
            return int(tok)
        except ValueError:
            pass
        try:
            return float(tok)
        except ValueError:
            return tok

    def block():
        nonlocal pos
        out = {}
        while pos < len(tokens) and tokens[pos] != "}":
            key = tokens[pos]
            pos += 1
            if tokens[pos] == "{":
                pos += 1
                out[key] = block()
                pos += 1  # the closing brace
            else:
                assert tokens[pos] == "=", (key, tokens[pos])
                out[key] = value(tokens[pos + 1])
                pos += 2
        return out

    return block()


@pytest.mark.parametrize("name", CONFS)
def test_shipped_conf_constructs_plugin(name):
    path = os.path.join(REF_CONFS, name)
    if not os.path.exists(path):
        pytest.skip("reference conf files not present")
    from neat_b200.loss import VolSDFLoss
    from neat_b200.model import VolSDFNetwork
    conf = parse_hocon(open(path).read())
    model = VolSDFNetwork(conf["model"])
    loss = VolSDFLoss(**conf["loss"])
    sd = model.state_dict()
    m = conf["model"]
    dims = m["implicit_network"]["dims"]
    assert sd["implicit_network.lin0.weight_v"].shape == (dims[0], 39)
    assert sd["implicit_network.lin%d.weight_v" % len(dims)].shape == (1 + m["feature_vector_size"], dims[-1])
    assert sd["rendering_network.lin0.weight_v"].shape[1] == 9 + 24 + m["feature_vector_size"]
    assert sd["attraction_network.lin4.weight_v"].shape[0] == 6
    assert sd["latents"].shape == (m["global_junctions"]["num_junctions"], m["global_junctions"]["dim_hidden"])
    assert float(sd["density.beta"]) == pytest.approx(m["density"]["params_init"]["beta"])
    assert model.dbscan_enabled == m["dbscan_enabled"] and model.use_median == m["use_median"]
    assert loss.eikonal_weight == conf["loss"]["eikonal_weight"] and loss.line_weight == conf["loss"]["line_weight"]
    # 3 tensors (weight_g, weight_v, bias) x (9 + 5 + 5) layers, density.beta, latents, 3 x (weight, bias) of ffn
    assert isinstance(model, torch.nn.Module) and len(list(model.parameters())) == 3 * 19 + 2 + 6
