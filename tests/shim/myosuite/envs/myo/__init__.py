from gym.envs.registration import register


def register_env_with_variants(id, entry_point, max_episode_steps, kwargs):
    """MyoSuite 1.2.3: the base env, plus sarcopenia / fatigue / reafferentation variants for ids that start with 'myo'
    (none of the reference's 'Custom...' ids do)."""
    register(id=id, entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=kwargs)
