import torch
from tests import flux_parity as FP
from oracle import flux_oracle as O, lokr_oracle as LO
cfg=FP.small_config(layers=2,single=2)
for linear_dim,alpha in ((10000,1),(8,4)):
    seed=0
    P={k:v.bfloat16().float() for k,v in O.init_flux_params(cfg,seed=seed).items()}
    shapes=O.flux_param_shapes(cfg)
    targets=[k[:-7] for k in shapes if k.endswith(".weight") and len(shapes[k])==2 and (".attn." in k or ".ff" in k)]
    fo=lambda n: 4 if (".ff." in n or ".ff_context." in n) else 10
    K={k:v.bfloat16().float() for k,v in LO.init_lokr_params({n:shapes[n+".weight"] for n in targets},linear_dim,fo,seed=seed+1,w2_std=0.02).items()}
    batch=FP.make_batch(2,16,16,64,cfg,seed=seed+2)
    g=torch.Generator().manual_seed(5)
    lat=batch["latent_batch"].float(); noise=torch.randn(lat.shape,generator=g).bfloat16().float(); sig=torch.rand(2,generator=g)
    noisy=O.flow_noisy_latents(lat.bfloat16(),noise.bfloat16(),sig).float()
    # effective weights as leaf tensors to get dW
    Pe={k:v.clone() for k,v in P.items()}
    Kg={k:v.clone().requires_grad_(True) for k,v in K.items()}
    O.LOKR={"linear_dim":linear_dim,"linear_alpha":alpha,"multiplier":1.0}
    Wl={}
    for n in targets:
        d=LO.lokr_delta(K,n,linear_dim,alpha)
        Pe[n+".weight"]=(P[n+".weight"]+d).requires_grad_(True)
    pred=O.flux_model_predict(Pe,cfg,noisy,sig*1000.0,batch["prompt_embeds"].float(),batch["add_text_embeds"].float(),1.0,None,1.0)
    loss=O.flow_loss(pred,O.flow_target(lat.bfloat16(),noise.bfloat16())); loss.backward()
    cos=torch.nn.functional.cosine_similarity
    worst=[]
    for n in targets:
        dW=Pe[n+".weight"].grad
        w1=K[n+".lokr_w1"]; w2=K[n+".lokr_w2"] if n+".lokr_w2" in K else K[n+".lokr_w2_a"]@K[n+".lokr_w2_b"]
        (a,c),(b,d)=w1.shape,w2.shape
        def contract(G):
            G=G.view(a,b,c,d); return torch.einsum("ajcl,jl->ac",G,w2), torch.einsum("ajcl,ac->jl",G,w1)
        e1,e2=contract(dW); r1,r2=contract(dW.bfloat16().float())
        worst.append((float(cos(e1.flatten(),r1.flatten(),dim=0)),float(cos(e2.flatten(),r2.flatten(),dim=0)),n,float(e1.norm()),float(dW.norm())))
    worst.sort()
    print(linear_dim,worst[:4])
