"""Display metrics with the reference's interface (tensorBNN/metrics.py).  The two forward passes they consume
come from the CUDA predict kernel; the reductions are torch ops on those device tensors and their results STAY on
the device (calculate() never synchronises) -- only display(), which prints, reads them back."""
import torch


class Metric(object):
    def __init__(self, scaleExp=False, mean=0, sd=1, *argv, **kwargs):
        self.scaleExp = scaleExp
        self.mean = mean
        self.sd = sd

    def _prep(self, predictionsTrain, predictionsValidate, realTrain, realValidate):
        pt = torch.as_tensor(predictionsTrain).t() * self.sd + self.mean
        pv = torch.as_tensor(predictionsValidate).t() * self.sd + self.mean
        rt = torch.as_tensor(realTrain).to(pt) * self.sd + self.mean
        rv = torch.as_tensor(realValidate).to(pv) * self.sd + self.mean
        if self.scaleExp:
            pt, pv, rt, rv = torch.exp(pt), torch.exp(pv), torch.exp(rt), torch.exp(rv)
        return pt, pv, rt.reshape(pt.shape), rv.reshape(pv.shape)

    def calculate(self, predictionsTrain, predictionValidate, realTrain, realValidate, *argv, **kwargs):
        pass

    def display(self):
        pass


class SquaredError(Metric):
    """Mean squared error (reference metrics.py:30-68; with scaleExp the reference exponentiates the
    training predictions but not the validation predictions -- reproduced)."""

    def calculate(self, predictionsTrain, predictionsValidate, realTrain, realValidate):
        pt = torch.as_tensor(predictionsTrain).t() * self.sd + self.mean
        pv = torch.as_tensor(predictionsValidate).t() * self.sd + self.mean
        rt = torch.as_tensor(realTrain).to(pt) * self.sd + self.mean
        rv = torch.as_tensor(realValidate).to(pv) * self.sd + self.mean
        if self.scaleExp:
            pt, rt, rv = torch.exp(pt), torch.exp(rt), torch.exp(rv)
        rt, rv = rt.reshape(pt.shape), rv.reshape(pv.shape)
        self.squaredErrorTrain = torch.mean((pt - rt) ** 2)
        self.squaredErrorValidate = torch.mean((pv - rv) ** 2)

    def display(self):
        print("training squared error{: 9.5f}".format(float(self.squaredErrorTrain)),
              "validation squared error{: 9.5f}".format(float(self.squaredErrorValidate)))


class PercentError(Metric):
    """Mean absolute percent error (reference metrics.py:70-108)."""

    def calculate(self, predictionsTrain, predictionsValidate, realTrain, realValidate):
        pt, pv, rt, rv = self._prep(predictionsTrain, predictionsValidate, realTrain, realValidate)
        self.percentErrorTrain = torch.mean(torch.abs((pt - rt) / rt) * 100)
        self.percentErrorValidate = torch.mean(torch.abs((pv - rv) / rv) * 100)

    def display(self):
        print("training percent error{: 7.3f}".format(float(self.percentErrorTrain)),
              "validation percent error{: 7.3f}".format(float(self.percentErrorValidate)))


class Accuracy(Metric):
    """Classification accuracy of rounded predictions (reference metrics.py:110-141)."""

    def calculate(self, predictionsTrain, predictionsValidate, realTrain, realValidate):
        pt, pv, rt, rv = self._prep(predictionsTrain, predictionsValidate, realTrain, realValidate)
        self.accuracyTrain = 1 - torch.mean(torch.abs(rt - torch.round(pt)))
        self.accuracyValidate = 1 - torch.mean(torch.abs(rv - torch.round(pv)))

    def display(self):
        print("training accuracy{: 9.5f}".format(float(self.accuracyTrain)),
              "validation accuracy{: 9.5f}".format(float(self.accuracyValidate)))
