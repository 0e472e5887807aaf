class AgentRenderVariant:  # name only (switch_env.py:15, distr_q.py:7)
    AGENT_SHOWS_OPTIONS = 3
