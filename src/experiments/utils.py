from peclr_b200.experiments_utils import (get_callbacks, get_checkpoints, get_general_args, get_model,  # noqa: F401
                                          prepare_name, restore_model, save_experiment_key, update_model_params,
                                          update_param, update_train_params)
