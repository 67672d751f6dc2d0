"""The "no defense" and "defense" workflows (drop-ins for recad/workflow/normal.py:162-225 and
recad/workflow/defense.py:175-303) driving the CUDA victims.

Same steps and the same call contract toward victims, attackers and datasets as the reference:
  1. train the victim `rec_epoch` epochs          (normal.py:170-177)
  2. train the attacker if it has a train_step      (normal.py:179-191)
  3. generate_fake -> dataset.inject_data -> victim.reset().I(dataset=fake) -> retrain (193-213)
  4. normal_evaluate(clean, attacked)               (normal.py:219-225)
The attacker (and, for "defense", the defender) is foreign code: any object with the reference's
attacker / defender interface; evaluation is the batched device evaluator of recad_b200.evaluate.
"""
import random

import numpy as np
import torch

from . import evaluate
from .config import SEED, WORKFLOW, get_logger, merge_config


class Normal:
    def __init__(self, **config):
        self.c = config
        self.attacker = config["attacker"].I(dataset=config["attack_data"]) if hasattr(config["attacker"], "I") else config["attacker"]
        self.victim = config["victim"].I(dataset=config["victim_data"])
        self.victim_data = config["victim_data"]
        self.logger = get_logger(__name__, level=self.c["logging_level"])
        self.results = None
        self.validation_log = []

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
        every = int(self.c.get("validate_every", 0) or 0)
        for e in range(config["epoch"]):
            loss = model.train_step(**self.info_describe(), progress_bar=None)
            out_des = model.output_describe()["train_step"]
            assert len(loss) == len(out_des), \
                f"The output describe is not aligned with the actual output of train_step for {model.model_name}"
            last = loss
            # optional per-epoch validation through the dataset's test-mode batches (implicit.py:461-476): the fused
            # full-rank + Recall/NDCG kernels make it cheap enough to run every epoch (off by default, as in the reference)
            if every and (e + 1) % every == 0 and hasattr(model, "full_rank") and hasattr(config.get("dataset"), "switch_mode"):
                res = evaluate.recall_ndcg_batches(model, config["dataset"], K=self.c.get("validate_k", 20), split="validate")
                self.validation_log.append({"epoch": e + 1, **res})
                self.logger.info(f"epoch {e + 1}: recall@{self.c.get('validate_k', 20)} {res['recall']:.5f} ndcg {res['ndcg']:.5f} "
                                 f"({res['n_users']} users)")
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


class Defense(Normal):
    """defense.py:17-303: the attack cycle of Normal (with the reference's re-seeding before the retrain),
    then  5. train the defender if it has a train_step, `fake_user_id = defender.defense_step()` (209-226)
          6. `victim_data.delete_data(...)` drops the flagged users, retrain from a fresh init, evaluate (228-303)."""

    def __init__(self, **config):
        super().__init__(**config)
        d = config["defender"]
        self.defender = d.I(dataset=config["defense_data"]) if hasattr(d, "I") else d
        self.results_after_defense = None

    @classmethod
    def from_config(cls, **kwargs):
        need = ("victim_data", "attack_data", "defense_data", "victim", "attacker", "defender")
        for k in need:                                                           # workflow/base.py:18-19
            if k not in kwargs:
                raise TypeError("Expect for user arguments [victim_data, attack_data, defense_data, victim, attacker, defender]")
        return cls(**merge_config(WORKFLOW["defense"], kwargs, need))

    def input_describe(self):
        return {"victim": "BaseVictim", "victim_data": "BaseData", "attacker": "BaseAttacker", "attack_data": "BaseData",
                "defender": "BaseDefender", "defense_data": "BaseData"}

    @staticmethod
    def random_seed_set():
        """defense.py:101-105"""
        random.seed(SEED)
        np.random.seed(SEED)
        torch.manual_seed(SEED)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(SEED)

    def execute(self):
        dev = self.c["device"]
        name = lambda m: getattr(m, "model_name", type(m).__name__)   # noqa: E731
        self.logger.info(f"Normal attacking, with dataset {self.victim_data.dataset_name}, victim model {self.victim.model_name}, "
                         f"attack model {name(self.attacker)}, defense model {name(self.defender)}, on device {dev}")
        self.victim = self.victim.to(dev)
        for role in ("attacker", "defender"):
            if hasattr(getattr(self, role), "to"):
                setattr(self, role, getattr(self, role).to(dev))
        self.logger.info("Step 1. training a recommender")
        self.normal_train(model=self.victim, epoch=self.c["rec_epoch"], dataset=self.victim_data)
        self.logger.info("Step 2. training a attacker")
        if "train_step" in self.attacker.input_describe():
            self.normal_train(model=self.attacker, epoch=self.c["attack_epoch"], dataset=self.c["attack_data"])
        fake_array = self.attacker.generate_fake(**self.info_describe())
        self.logger.info(f"Step 3. injecting fake data({tuple(fake_array.shape)}) and re-train the recommender")
        fake_dataset = self.victim_data.inject_data("explicit", fake_array, filter_num=self.c["filter_num"])
        self.logger.info("Step 4. Retraining a recommender")
        self.random_seed_set()
        fake_victim = self.victim.reset().I(dataset=fake_dataset).to(dev)
        self.normal_train(model=fake_victim, epoch=self.c["rec_epoch"], dataset=fake_dataset)
        self.fake_victim, self.fake_dataset = fake_victim, fake_dataset
        self.results = self.normal_evaluate(self.victim, fake_victim, self.victim_data, self.c["target_id_list"],
                                            topks=self.c["topks"])
        self.logger.info("Step 5. training a defender. ")
        if "train_step" in self.defender.input_describe():
            self.normal_train(model=self.defender, epoch=self.c["defense_epoch"], dataset=fake_dataset)
        fake_user_id = self.defender.defense_step()
        self.logger.info(f"Step 6. Delete fake data(len = {len(fake_user_id)}) and re-train the recommender")
        cleaned = self.victim_data.delete_data("explicit", fake_user_id, fake_array, filter_num=self.c["filter_num"])
        self.random_seed_set()
        defended = self.victim.reset().I(dataset=cleaned).to(dev)
        self.normal_train(model=defended, epoch=self.c["rec_epoch"], dataset=cleaned)
        self.defended_victim, self.cleaned_dataset, self.flagged_users = defended, cleaned, list(fake_user_id)
        self.results_after_defense = self.normal_evaluate(self.victim, defended, self.victim_data, self.c["target_id_list"],
                                                          topks=self.c["topks"])
        return self.results_after_defense


factories = {"no defense": Normal, "defense": Defense}     # recad/workflow/__init__.py:4-7


def from_config(name, **kwargs):
    return factories[name].from_config(**kwargs)
