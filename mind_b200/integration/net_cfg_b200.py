"""Network configuration module for MIND's planner (`planner_cfg["network_config"]`, planners/mind/planner.py:42-49):
the reference's own hyper-parameters (planners/mind/configs/networks/net_cfg.py) with the network class swapped for the
B200 predictor.  This string is the whole drop-in: `MINDPlanner.init_network` imports the class, loads the shipped
checkpoint's 328-key state_dict into it, `.to(device)`, `.eval()` -- unchanged."""
from planners.mind.configs.networks.net_cfg import NetCfg as _ReferenceNetCfg


class NetCfg(_ReferenceNetCfg):
    def get_net_cfg(self):
        cfg = super().get_net_cfg()
        cfg["network"] = "mind_b200.predictor:ScenePredNetB200"
        return cfg
