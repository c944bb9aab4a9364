def make_roi_box_loss_evaluator(cfg):
    return None  # training-only
