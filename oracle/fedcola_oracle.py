"""TEST INFRASTRUCTURE ONLY — CPU restatement of FedCola's federated-round hot path.

This file is the *oracle*: a plain numpy / torch-CPU-fp32 restatement of the reference algorithm.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it, and only as the checker / the CPU baseline — never as the product path.  The product
(`fedcola_b200/`) fails loudly when its CUDA library is missing; it never falls back to this file.

Parity pin: every function below is checked against the UNMODIFIED reference imported under
`oracle/ref_shim.py` (tests/test_oracle_vs_reference.py, runs where /root/reference exists) and against
golden vectors generated from the reference by `oracle/make_golden.py` (tests/golden/*.npz, travel to
the GPU box).  Two third-party pieces the reference depends on are absent from /root/reference and
from this image — timm==0.9.12 `DropPath` and torchmultimodal `ContrastiveLossWithTemperature`
(unpinned) — and are restated from their published formulas: **parity unpinned** for those two
(corroborated against torchvision.ops.stochastic_depth and transformers' clip_loss).

Reference map (file:line under /root/reference):
  get_name_type / get_first_number / get_name_modality   src/server/fedavgserver.py:94-115
  init_param_scope                                       src/server/fedavgserver.py:183-238
  sample_clients                                         src/server/fedavgserver.py:282-312
  coefficients                                           src/server/fedavgserver.py:601-653
  aggregate_lerp                                         src/server/fedavgserver.py:597,656-666
  upload_merge                                           src/client/fedavgclient.py:158-184
  aux_refresh                                            src/server/fedavgserver.py:821-845
  mat_forward & friends                                  src/models/mome.py:42-60,100-228,578-659,881-922
  client_update                                          src/client/fedavgclient.py:55-116, fedproxclient.py:17-88
"""
import math
import random
import re

import numpy as np
import torch
import torch.nn.functional as F

AUX_LAYERS_ALL = ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")


# --------------------------------------------------------------------------------------------------
# name bookkeeping (fedavgserver.py:94-115, 183-238)
# --------------------------------------------------------------------------------------------------
def get_name_type(name):
    if "embeddings" in name:
        return "embedding"
    elif "attention" in name or "attn" in name:
        return "attn"
    elif "blocks" in name:
        return "blocks"
    elif "mlp" in name:      # unreachable for blockses.* keys, kept as in the reference
        return "mlp"
    return "task"


def get_first_number(string):
    m = re.search(r"\d+", string)
    return int(m.group()) if m else None


def get_name_modality(name, modalities):
    idx = get_first_number(name)
    return modalities[idx] if idx is not None else None


def init_param_scope(names, shared_param, share_scope):
    scope = {}
    for name in names:
        t = get_name_type(name)
        if shared_param == "none":
            scope[name] = "dataset"
        elif shared_param == "attn":
            scope[name] = share_scope if t == "attn" else "dataset"
        elif shared_param == "blocks":
            scope[name] = share_scope if t == "blocks" else "dataset"
        elif shared_param == "mlp":
            scope[name] = share_scope if t == "mlp" else "dataset"
    return scope


# --------------------------------------------------------------------------------------------------
# client sampling (fedavgserver.py:282-312) — python `random`, bit-exact
# --------------------------------------------------------------------------------------------------
def sample_clients(datasets, Cs, client_dataset_of, equal_sampled, C=None, K=None):
    """client_dataset_of: list, index = client id -> dataset name.  Uses the global `random` state."""
    if equal_sampled:
        out = []
        for ds in datasets:
            ids = [i for i, d in enumerate(client_dataset_of) if d == ds]
            n = max(int(Cs[ds] * len(ids)), 1)
            out += sorted(random.sample(ids, n))
        return sorted(out)
    n = max(int(C * K), 1)
    return sorted(random.sample([i for i in range(K)], n))


# --------------------------------------------------------------------------------------------------
# mixing coefficients (fedavgserver.py:601-653)
# --------------------------------------------------------------------------------------------------
def coefficients(param_names, param_scope, clients, updated_sizes, g_dataset, g_modality, g_task,
                 out_modality_scale, args_modalities, share_scope_flag, compensation, fedavg=False):
    """clients: {id: dict(dataset=, modality=, task=)}; updated_sizes: {id: n} in the reference's dict
    order.  Returns {param: {id: python float}} exactly as the reference builds it."""
    coefs = {}
    ids = list(updated_sizes.keys())
    for p in param_names:
        num = {}
        old_sum = sum(updated_sizes.values())
        pm = None if fedavg else get_name_modality(p, args_modalities)
        sc = param_scope[p]
        for k, n in updated_sizes.items():
            c = clients[k]
            if sc == "all":
                num[k] = n
            elif sc == "dataset":
                num[k] = n if c["dataset"] == g_dataset else 0
            elif sc == "task":
                num[k] = n if c["task"] == g_task else 0
            elif sc == "modality":
                if fedavg:
                    num[k] = n if c["modality"] == g_modality else 0
                else:
                    num[k] = n if (c["modality"] in g_modality or g_modality in c["modality"]) else 0
            elif sc == "modality_exact" and not fedavg:
                num[k] = n if (c["modality"] == pm or pm in c["modality"]) else 0
            if not fedavg and c["modality"] != g_modality and out_modality_scale != 1:
                old_sum -= num[k]
                num[k] *= out_modality_scale
                old_sum += num[k]
        if compensation and not fedavg:
            if share_scope_flag == "all":
                coefs[p] = {k: float(v / old_sum) for k, v in num.items()}
            elif share_scope_flag == "modality":
                comp = sum(s for i, s in updated_sizes.items()
                           if clients[i]["modality"] in g_modality or g_modality in clients[i]["modality"])
                coefs[p] = {k: float(v / comp) if comp != 0 else 0 for k, v in num.items()}
            elif share_scope_flag == "modality_exact":
                if pm:
                    last = ids[-1]   # the reference's stale loop variable `identifier` (fedavgserver.py:648)
                    comp = sum(s for i, s in updated_sizes.items()
                               if clients[i]["modality"] == pm or pm in clients[last]["modality"])
                else:
                    comp = sum(s for i, s in updated_sizes.items()
                               if clients[i]["modality"] in g_modality or g_modality in clients[i]["modality"])
                coefs[p] = {k: float(v / comp) if comp != 0 else 0 for k, v in num.items()}
            else:
                # share_scope == 'dataset' with --compensation: the reference leaves coefficients[p] unset
                # (KeyError at :657).  Mirror the error.
                raise KeyError(p)
        else:
            tot = sum(num.values())
            coefs[p] = {k: float(v / tot) if tot != 0 else 0 for k, v in num.items()}
    return coefs


# --------------------------------------------------------------------------------------------------
# upload merge + sequential-lerp aggregation (fedavgclient.py:158-184; fedavgserver.py:656-666)
# --------------------------------------------------------------------------------------------------
def aux_layer_names(aux_attn_only=False, aux_mlp_only=False):
    if aux_attn_only:
        if aux_mlp_only:
            raise ValueError("Both aux_attn_only and aux_mlp_only cannot be True.")
        return ("attn.qkv", "attn.proj")
    if aux_mlp_only:
        return ("mlp.fc1", "mlp.fc2")
    return AUX_LAYERS_ALL


def upload_merge(sd, with_aux, modality, aux_attn_only=False, aux_mlp_only=False):
    """sd: {key: np.float32 array}.  Returns what `FedavgClient.upload()` hands to the server."""
    if not (with_aux and modality != "img+txt"):
        return dict(sd)
    names = aux_layer_names(aux_attn_only, aux_mlp_only)
    new = {k: v.copy() for k, v in sd.items()}
    for k, v in sd.items():
        if any(n in k for n in names) and "aux" not in k and "weight" in k:
            a = new[k.replace("weight", "aux_weight")]
            s = new[k.replace("weight", "cross_modal_scale")]
            new[k] = (v + (a * s).astype(np.float32)).astype(np.float32)     # mul then add, fp32 roundings
    for k in sd:
        if "aux" in k or "cross_modal_scale" in k:
            new.pop(k)
    return new


def aggregate_lerp(final_sd, uploads, coefs, ids):
    """final_sd: {param: np.float32 array} = old global's required_params (modified in place and
    returned).  uploads: {id: {param: array}}.  Exactly `f += ((l - f) * c32)` per sampled id in order."""
    for k in ids:
        local = uploads[k]
        for p in coefs:
            c = coefs[p][k]
            if p not in local or c == 0:
                continue
            f = final_sd[p]
            c32 = np.float32(c)
            t = (local[p].astype(np.float32) - f).astype(np.float32)
            t = (t * c32).astype(np.float32)
            final_sd[p] = (f + t).astype(np.float32)
    return final_sd


def closed_form_weights(coefs_p, ids, present):
    """fp64 closed form of the sequential lerp for one param: returns (w_global, {id: w_k}) with
    result = w_global*g + sum_k w_k*l_k.  `present`: ids that upload this key."""
    w = {}
    wg = 1.0
    for k in ids:                       # ascending
        c = float(np.float32(coefs_p[k])) if (k in present and coefs_p[k] != 0) else 0.0
        for j in w:
            w[j] *= (1.0 - c)
        wg *= (1.0 - c)
        w[k] = c
    return wg, w


def aux_refresh_map(aux_keys, own_idx):
    """{aux key in this uni-modal global: key in the other-modality global}  (fedavgserver.py:821-845)."""
    other = 1 - own_idx
    return {k: k.replace("aux_", "").replace(f"blockses.{own_idx}", f"blockses.{other}") for k in aux_keys}


# --------------------------------------------------------------------------------------------------
# ModalityAgnosticTransformer forward, functional, torch fp32 (mome.py)
# --------------------------------------------------------------------------------------------------
def _linear(p, name, x):
    w = p[name + ".weight"]
    if (name + ".aux_weight") in p:                      # CrossModalReparamLinear.forward, mome.py:58-60
        w = w + p[name + ".cross_modal_scale"] * p[name + ".aux_weight"]
    return F.linear(x, w, p[name + ".bias"])


def attention_forward(p, pre, x, num_heads):             # mome.py:150-168 (no mask ever applied)
    B, N, C = x.shape
    hd = C // num_heads
    qkv = _linear(p, pre + "qkv", x).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    q = q * (hd ** -0.5)
    attn = q.float() @ k.float().transpose(-2, -1)
    attn = attn.softmax(dim=-1).type_as(x)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return _linear(p, pre + "proj", x)


def block_forward(p, pre, x, num_heads, dp1=None, dp2=None):   # mome.py:225-228
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
    h = attention_forward(p, pre + "attn.", h, num_heads)
    x = x + (h if dp1 is None else h * dp1)
    h = F.layer_norm(x, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    h = _linear(p, pre + "mlp.fc1", h)
    h = F.gelu(h)                                         # nn.GELU() = exact erf
    h = _linear(p, pre + "mlp.fc2", h)
    return x + (h if dp2 is None else h * dp2)


def image_embed(p, pre, img):                             # mome.py:597-611 + PatchEmbed :260-266
    if img.shape[1] == 1:
        img = img.repeat(1, 3, 1, 1)                      # mome.py:893-894
    x = F.conv2d(img, p[pre + "embed.proj.weight"], p[pre + "embed.proj.bias"], stride=16)
    x = x.flatten(2).transpose(1, 2)
    cls = p[pre + "cls_token"].expand(x.shape[0], -1, -1)
    x = torch.cat((cls, x), dim=1)
    return x + p[pre + "pos_embed"]


def text_embed(p, pre, ids):                              # mome.py:632-639; HF BertEmbeddings, eps 1e-12
    L = ids.shape[1]
    t = pre + "text_embeddings."
    # nn.Embedding(vocab, d, padding_idx=pad_token_id=0): row 0 is looked up normally but never gets gradient
    x = F.embedding(ids, p[t + "word_embeddings.weight"], padding_idx=0) + p[t + "token_type_embeddings.weight"][0] \
        + p[t + "position_embeddings.weight"][:L]
    C = x.shape[-1]
    return F.layer_norm(x, (C,), p[t + "LayerNorm.weight"], p[t + "LayerNorm.bias"], 1e-12)


def mat_forward(p, x, modalities, num_heads, depth, feat_out=False, drop_path_rate=0.0, training=True):
    """p: {state_dict key: tensor}; x = [img|None, ids|None]; returns [out0|None, out1|None]  (mome.py:881-922).
    DropPath masks are drawn per modality encoder in block order, attn branch then mlp branch."""
    embeds = []
    for i, m in enumerate(modalities):
        if m is None:
            embeds.append(None)
        elif m == "img":
            embeds.append(image_embed(p, f"embeddings.{i}.", x[i]))
        else:
            embeds.append(text_embed(p, f"embeddings.{i}.", x[i]))
    outs = [None, None]
    C = p["norm.weight"].shape[0]
    for i, m in enumerate(modalities):
        if m is None:
            continue
        h = embeds[i]
        if drop_path_rate > 0 and training:
            # the reference draws inside each Block.forward → block j of encoder i: dp1 then dp2
            dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]
        for j in range(depth):
            dp1 = dp2 = None
            if drop_path_rate > 0 and training and dpr[j] > 0:
                keep = 1 - dpr[j]
                pre = f"blockses.{i}.{j}."
                # attn branch mask is drawn after attn() ran, mlp branch mask after mlp() ran — neither
                # consumes RNG, so drawing them just-in-time here is the same sequence.
                dp1 = torch.empty((h.shape[0], 1, 1)).bernoulli_(keep).div_(keep)
                hh = F.layer_norm(h, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
                hh = attention_forward(p, pre + "attn.", hh, num_heads)
                h = h + hh * dp1
                dp2 = torch.empty((h.shape[0], 1, 1)).bernoulli_(keep).div_(keep)
                hh = F.layer_norm(h, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
                hh = _linear(p, pre + "mlp.fc2", F.gelu(_linear(p, pre + "mlp.fc1", hh)))
                h = h + hh * dp2
            else:
                h = block_forward(p, f"blockses.{i}.{j}.", h, num_heads)
        h = F.layer_norm(h, (C,), p["norm.weight"], p["norm.bias"], 1e-6)
        if feat_out:
            c = h[:, 0]
            outs[i] = c / c.norm(dim=-1, keepdim=True)
        elif (f"heads.{i}.head.weight") in p:
            outs[i] = F.linear(h[:, 0], p[f"heads.{i}.head.weight"], p[f"heads.{i}.head.bias"])
        else:                                             # RetrievalHead (mome.py:651-659)
            c = h[:, 0]
            outs[i] = c / c.norm(dim=-1, keepdim=True)
    return outs


LOGIT_SCALE = math.log(1 / 0.07)


def contrastive_loss(a, b):
    """torchmultimodal ContrastiveLossWithTemperature, fresh object every step → constant τ."""
    lab = torch.arange(a.size(0))
    tt = torch.exp(torch.tensor(LOGIT_SCALE, dtype=torch.float32))
    return (F.cross_entropy(a @ b.t() * tt, lab) + F.cross_entropy(b @ a.t() * tt, lab)) / 2


def client_loss(p, batch, modality, modalities, num_heads, depth, drop_path_rate=0.0):
    """One forward + loss as in FedavgClient.update (fedavgclient.py:81-95). Returns (loss, outputs)."""
    if modality == "img":
        x, y = batch[0], batch[1]
        out = mat_forward(p, [x, None], modalities, num_heads, depth, drop_path_rate=drop_path_rate)[0]
        return F.cross_entropy(out, y), out
    if modality == "txt":
        x, y = batch[0], batch[1]
        out = mat_forward(p, [None, x], modalities, num_heads, depth, drop_path_rate=drop_path_rate)[1]
        return F.cross_entropy(out, y), out
    img, ids = batch[0], batch[1]
    outs = mat_forward(p, [img, ids], modalities, num_heads, depth, feat_out=True, drop_path_rate=drop_path_rate)
    return contrastive_loss(*outs), outs[0]


def client_update(params, requires_grad, batches, modality, modalities, num_heads, depth, E=1,
                  optimizer="AdamW", lr=1e-4, weight_decay=0.0, momentum=0.0, nesterov=False,
                  max_grad_norm=0.0, mu=None, drop_path_rate=0.0, n_total=None):
    """Local training loop restated (fedavgclient.py:55-116; prox term fedproxclient.py:64-67 when mu is
    given).  params: {key: tensor} (aliases may map two keys to the same tensor), updated in place.
    Returns (per-step losses, {epoch: mean loss})."""
    uniq, seen = [], set()
    for k, v in params.items():
        if requires_grad.get(k, True) and id(v) not in seen:
            seen.add(id(v))
            v.requires_grad_(True)
            uniq.append((k, v))
    plist = [v for _, v in uniq]
    if optimizer == "AdamW":
        opt = torch.optim.AdamW(plist, lr=lr, weight_decay=weight_decay)
    elif optimizer == "SGD":
        opt = torch.optim.SGD(plist, lr=lr, momentum=momentum, nesterov=nesterov, weight_decay=weight_decay)
    else:
        raise NotImplementedError(optimizer)
    if mu is not None:
        frozen = [v.detach().clone() for v in plist]
    losses, results = [], {}
    for e in range(E):
        run, cnt = 0.0, 0
        for batch in batches:
            opt.zero_grad()
            loss, out = client_loss(params, batch, modality, modalities, num_heads, depth, drop_path_rate)
            if mu is not None:
                prox = 0.0
                for v, g in zip(plist, frozen):
                    prox = prox + (v - g).norm(2)
                loss = loss + mu * (0.5 * prox)
            loss.backward()
            if max_grad_norm > 0:
                torch.nn.utils.clip_grad_norm_(plist, max_grad_norm)
            opt.step()
            losses.append(loss.item())
            run += loss.item() * len(out)
            cnt += len(out)
        results[e + 1] = {"loss": run / (n_total if n_total else cnt)}
    for v in plist:
        v.requires_grad_(False)
    return losses, results
