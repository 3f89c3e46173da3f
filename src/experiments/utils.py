from peclr_b200.experiments_utils import (get_callbacks, get_general_args, get_model, update_model_params,  # noqa: F401
                                          update_param, update_train_params)
