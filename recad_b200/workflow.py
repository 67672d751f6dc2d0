"""The "no defense" workflow (drop-in for recad/workflow/normal.py:162-225) driving the CUDA victims.

Same steps and the same call contract toward victims, attackers and datasets as the reference:
  1. train the victim `rec_epoch` epochs          (normal.py:170-177)
  2. train the attacker if it has a train_step      (normal.py:179-191)
  3. generate_fake -> dataset.inject_data -> victim.reset().I(dataset=fake) -> retrain (193-213)
  4. normal_evaluate(clean, attacked)               (normal.py:219-225)
The attacker is foreign code (any object with the reference's attacker interface); evaluation is
the batched device evaluator of recad_b200.evaluate.
"""
import os

from . import evaluate
from .config import WORKFLOW, get_logger, merge_config


class Normal:
    def __init__(self, **config):
        self.c = config
        self.attacker = config["attacker"].I(dataset=config["attack_data"]) if hasattr(config["attacker"], "I") else config["attacker"]
        self.victim = config["victim"].I(dataset=config["victim_data"])
        self.victim_data = config["victim_data"]
        self.logger = get_logger(__name__, level=self.c["logging_level"])
        self.results = None

    @classmethod
    def from_config(cls, **kwargs):
        for need in ("victim_data", "attack_data", "victim", "attacker"):      # workflow/base.py:18-19
            if need not in kwargs:
                raise TypeError("Expect for user arguments [victim_data, attack_data, victim, attacker]")
        cfg = merge_config(WORKFLOW["no defense"], kwargs, ("victim_data", "attack_data", "victim", "attacker"))
        return cls(**cfg)

    def info_describe(self):
        return {"target_id_list": self.c["target_id_list"], "input_describe": self.input_describe()}

    def input_describe(self):
        return {"victim": "BaseVictim", "victim_data": "BaseData", "attacker": "BaseAttacker", "attack_data": "BaseData"}

    def normal_train(self, **config):
        """normal.py:95-109: one train_step per epoch; its tuple must match output_describe()."""
        model = config["model"]
        last = None
        for _ in range(config["epoch"]):
            loss = model.train_step(**self.info_describe(), progress_bar=None)
            out_des = model.output_describe()["train_step"]
            assert len(loss) == len(out_des), \
                f"The output describe is not aligned with the actual output of train_step for {model.model_name}"
            last = loss
        return last

    def normal_evaluate(self, model, model_fake, dataset, target_id_list, topks):
        return evaluate.normal_evaluate(model, model_fake, dataset, target_id_list, topks,
                                        verbose=self.c.get("verbose", True))

    def execute(self):
        dev = self.c["device"]
        self.logger.info(f"Normal attacking, with dataset {self.victim_data.dataset_name}, victim model "
                         f"{self.victim.model_name}, attack model {getattr(self.attacker, 'model_name', type(self.attacker).__name__)}, "
                         f"on device {dev}")
        self.victim = self.victim.to(dev)
        if hasattr(self.attacker, "to"):
            self.attacker = self.attacker.to(dev)
        self.logger.info("Step 1. training a recommender")
        self.normal_train(model=self.victim, epoch=self.c["rec_epoch"], dataset=self.victim_data)
        self.logger.info("Step 2. training a attacker")
        if "train_step" in self.attacker.input_describe():
            self.normal_train(model=self.attacker, epoch=self.c["attack_epoch"], dataset=self.c["attack_data"])
        fake_array = self.attacker.generate_fake(**self.info_describe())
        self.logger.info(f"Step 3. injecting fake data({tuple(fake_array.shape)}) and re-train the recommender")
        fake_dataset = self.victim_data.inject_data("explicit", fake_array, filter_num=self.c["filter_num"])
        fake_victim = self.victim.reset().I(dataset=fake_dataset)
        fake_victim = fake_victim.to(dev)
        self.normal_train(model=fake_victim, epoch=self.c["rec_epoch"], dataset=fake_dataset)
        self.fake_victim, self.fake_dataset = fake_victim, fake_dataset
        self.results = self.normal_evaluate(self.victim, fake_victim, self.victim_data, self.c["target_id_list"],
                                            topks=self.c["topks"])
        return self.results


factories = {"no defense": Normal}     # recad/workflow/__init__.py:4


def from_config(name, **kwargs):
    return factories[name].from_config(**kwargs)
