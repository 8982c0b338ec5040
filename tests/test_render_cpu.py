"""Render helpers against strings recorded from the unmodified reference
(tests/golden/render_n3.json, made by tests/golden/make_render_golden.py)."""
import json
import os

from skyjo_rl_b200.env import render_action_explainer, render_actions

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_n3.json")))


def test_action_explainer_and_legend_match_reference():
    assert [render_action_explainer(a) for a in range(26)] == GOLD["explainer"]     # skyjo.py:566-590
    assert render_actions() == GOLD["render_actions"]                               # skyjo.py:592-602
