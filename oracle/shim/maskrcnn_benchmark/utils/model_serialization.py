def load_state_dict(model, loaded_state_dict):
    """Suffix-matching loader (upstream): strip 'module.', match each model key to the
    loaded key that ends with it."""
    loaded = {k[7:] if k.startswith("module.") else k: v for k, v in loaded_state_dict.items()}
    own = model.state_dict()
    for k in own:
        cands = [lk for lk in loaded if lk == k or lk.endswith("." + k) or k.endswith("." + lk)]
        if cands:
            own[k] = loaded[max(cands, key=len)]
    model.load_state_dict(own)
