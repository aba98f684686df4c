"""Environment-name validation (reference ``olympic_mujoco/utils/checks.py:3-76``): same conditions, same
exception type (ValueError); messages are shortened."""


def check_validity_task_mode_dataset(env_name, task=None, mode=None, dataset_type=None, valid_tasks=None,
                                     valid_modes=None, valid_dataset_types=None, non_combineable=None):
    parts = [p for p, v in (("<task>", task), ("<mode>", mode), ("<dataset_type>", dataset_type)) if v is not None]
    hint = f"\n\nThe general structure for calling the environment {env_name} is:\n{env_name}." + ".".join(parts)
    if task is not None and task not in valid_tasks:
        raise ValueError(f'Task "{task}" does not exit in the environment {env_name}. Please, choose from '
                         f"{valid_tasks}. {hint}")
    if mode is not None and mode not in valid_modes:
        raise ValueError(f'Mode "{mode}" does not exit in the environment {env_name}. Please, choose from '
                         f"{valid_modes}. {hint}")
    if dataset_type is not None and dataset_type not in valid_dataset_types:
        raise ValueError(f'Dataset type "{dataset_type}" does not exit in the environment {env_name}. '
                         f"Please, choose from {valid_dataset_types}. {hint}")
    if non_combineable is not None:
        for bad_t, bad_m, bad_dt in non_combineable:
            if (task == bad_t or bad_t is None) and (mode == bad_m or bad_m is None) \
                    and (dataset_type == bad_dt or bad_dt is None):
                raise ValueError(f'Task "{task}", mode "{mode}" and dataset type "{dataset_type}" are not combineable '
                                 f"for the environment {env_name}. {hint}")
