"""Task instruction strings of the labeler (reference: arp_dt/data_procgen.py:281-317)."""

_GOAL = {
    "coinrun": "the goal is to collect the coin.",
    "coinrun_aisc": "the goal is to collect the coin.",
    "maze": "navigate a maze to collect the yellow cheese.",
    "maze_aisc": "navigate a maze to collect the yellow cheese.",
    "maze_yellowline": "navigate a maze to collect the yellow line.",
    "maze_redline_yellowgem": "navigate a maze to collect the red line.",
}

# multi-instruction (positive / negative) prompts for Maze II / III style labeling
# (reference: arp_dt/assets/procgen_instruct.py:72-105, consumed by envs/vl_reward.py with a list)
POS_NEG = {
    "coinrun": ["The goal is to collect the coin.", "The agent must navigate to the far right wall."],
    "coinrun_aisc": ["The goal is to collect the coin.", "The agent must navigate to the far right wall."],
    "maze": ["The agent must navigate a maze to find the yellow cheese.", "The agent navigate to the top right."],
    "maze_aisc": ["The agent must navigate a maze to find the yellow cheese.", "The agent navigate to the top right."],
    "maze_yellowline": ["The agent must navigate a maze to find the line.", "The agent navigate to the yellow object."],
    "maze_redline_yellowgem": ["The agent must navigate a maze to find the line.",
                               "The agent navigate to the yellow object."],
    "maze_yellowstar_redgem": ["The agent must navigate a maze to find the yellow objects.",
                               "The agent must dodge the red objects."],
}


def get_clip_instruct(task: str):
    """data_procgen.py:281-293 — returns None for an unknown task, exactly like the reference's if-chain."""
    return _GOAL.get(task)


def get_clip_special_instruct(env_name: str, inst_type: str) -> str:
    """data_procgen.py:296-317."""
    if inst_type == "random1":
        return "His voice echoed through the empty hallway."
    if inst_type == "random2":
        return "NeurIPS 2023 will be held again at the at the New Orleans Ernest N. Morial Convention Center."
    if inst_type == "misinfo":
        if "coinrun" in env_name:
            return "The agent must go to the far right of the level."
        if env_name == "maze_aisc":
            return "navigate a maze to reacth to the top right corner."
        if env_name == "maze_yellowline":
            return "navigate a maze to collect yellow gem."
    elif "coinrun" in env_name:
        special = {"misinfo2": "The goal is to collect the red strawberry.", "misinfo3": "The goal is to reach the saw.",
                   "misinfo4": "The goal is to jump as high as you can."}
        if inst_type in special:
            return special[inst_type]
    raise ValueError("You must pass any condition.")
